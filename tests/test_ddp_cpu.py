"""Host-side data-parallel logic on CPU: gradient buckets partition the flat gradient buffer in backward-completion
order, and a world_size-2 `gloo` run of the bucketed all-reduce reproduces the gradient of the global batch
(per-replica BatchNorm statistics, as in the GPU path).  The oracle is used here only as the per-rank gradient
producer / checker."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _small_net():
    from dfcsa.modules import UNetDFCSARes
    torch.manual_seed(0)
    return UNetDFCSARes(3, 1, [4, 8, 16, 32], pool_size=4, ablation_on_qk_channels=4)


def test_buckets_partition_the_flat_gradient_buffer_in_backward_order():
    from dfcsa.ddp import FLAT_ALIGN, bucket_modules, bucket_ranges, flat_offsets
    net = _small_net()
    ranges = bucket_ranges(net)            # asserts: every parameter in exactly one bucket
    assert len(ranges) == 9 and all(len(r) == 1 for r in ranges)
    offs, total = flat_offsets(list(net.parameters()))
    assert all(lo % FLAT_ALIGN == 0 for lo, _ in offs.values())              # 128-byte aligned tensors
    runs = [r[0] for r in ranges]
    for lo, hi in offs.values():                                            # each tensor inside exactly one run
        assert sum(1 for a, b in runs if a <= lo and hi <= b) == 1
    assert all(a2 >= b1 for (a1, b1), (a2, b2) in zip(sorted(runs), sorted(runs)[1:]))   # runs do not overlap
    # reverse execution order: the first bucket ends with the last tensor (final_conv.bias), the last one starts at 0
    assert ranges[0][0][1] == max(hi for _, hi in offs.values()) and ranges[-1][0][0] == 0
    assert [type(m).__name__ for m in bucket_modules(net)[4]] == ["ConvTranspose2d"]      # {up4} precedes the bottleneck
    # the bottleneck bucket is the largest one (42 % of the bytes at full width)
    sizes = [sum(hi - lo for lo, hi in runs_) for runs_ in ranges]
    assert sizes.index(max(sizes)) == 5


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard_grads(sd, names, img, mask):
    from oracle import dfcsa_oracle as O
    ps = {k: (v.clone().requires_grad_(True) if k in names else v.clone()) for k, v in sd.items()}
    logits = O.unet_forward(img, ps, 4, training=True)
    loss = O.calculate_metrics(torch.sigmoid(logits), mask, "bce_dice", {})["loss"]
    return torch.autograd.grad(loss, [ps[k] for k in names])


def _worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "dfc-sa-unet_b200")]
    from dfcsa.ddp import BucketReducer, bucket_ranges, flat_offsets
    from oracle import dfcsa_oracle as O
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    net = _small_net()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    names = [n for n, _ in net.named_parameters()]
    img, mask = O.synthetic_batch(2 * world, 32, 32, seed=5)
    g = _shard_grads(sd, names, img[2 * rank:2 * rank + 2], mask[2 * rank:2 * rank + 2])
    offs, total = flat_offsets(list(net.parameters()))
    flat = torch.zeros(total)
    for p, t in zip(net.parameters(), g):            # the layout FusedSGD.flat_grad uses
        flat[offs[p][0]:offs[p][1]] = t.reshape(-1)
    red = BucketReducer(flat, bucket_ranges(net))
    calls = []
    for k in range(len(red.ranges)):        # the order net_backward's after_stage(k) fires in
        red.reduce(k)
        calls.append(k)
    red.finish()
    flat *= 1.0 / world                     # what the fused SGD kernel's grad_scale does
    dense = torch.cat([flat[offs[p][0]:offs[p][1]] for p in net.parameters()])
    torch.save({"flat": dense, "calls": calls}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_gloo_world2_bucketed_allreduce_matches_mean_of_shard_gradients(tmp_path):
    from oracle import dfcsa_oracle as O
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    assert torch.equal(outs[0]["flat"], outs[1]["flat"])            # every rank ends with the same gradient
    net = _small_net()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    names = [n for n, _ in net.named_parameters()]
    img, mask = O.synthetic_batch(2 * world, 32, 32, seed=5)
    ref = None
    for r in range(world):
        g = torch.cat([t.reshape(-1) for t in _shard_grads(sd, names, img[2 * r:2 * r + 2], mask[2 * r:2 * r + 2])])
        ref = g if ref is None else ref + g
    ref /= world
    assert torch.allclose(outs[0]["flat"], ref, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("n,world,bs", [(127, 2, 64), (128, 2, 64), (129, 2, 64), (10, 4, 3), (7, 8, 2), (1000, 8, 64), (5, 2, 8),
                                        (64, 3, 7), (1, 2, 1)])
@pytest.mark.parametrize("drop_last", [True, False])
def test_train_sharding_gives_every_rank_the_same_batches(n, world, bs, drop_last):
    """Every train step issues per-bucket all-reduces, so all ranks must run the same number of steps with the same
    batch sizes whatever n / world / batch size are (the order[rank::world] slicing of round 1 could leave low ranks one
    step ahead and hang the job: n=127, world=2, bs=64)."""
    from dfcsa.data_loader import GpuLoader
    order = list(range(n))
    per_rank = [GpuLoader.rank_batches(order, bs, r, world, drop_last) for r in range(world)]
    shapes = [[len(b) for b in br] for br in per_rank]
    assert all(s == shapes[0] for s in shapes), shapes
    seen = [i for br in per_rank for b in br for i in b]
    if drop_last:
        assert all(len(b) == bs for br in per_rank for b in br)
        assert len(seen) == len(set(seen)) == n // (world * bs) * (world * bs)          # no duplicates, whole global batches
    else:
        assert set(seen) == set(order)                                                  # every sample is visited
        assert len(seen) == (n + world - 1) // world * world                            # padded by wrapping around
    # __len__ agrees with the iteration
    ld = GpuLoader(order, bs, (8, 8), augment=False, shuffle=False, drop_last=drop_last, rank=world - 1, world=world)
    assert len(ld) == len(per_rank[-1])


@pytest.mark.parametrize("n,world,bs", [(10, 2, 3), (9, 4, 2), (3, 4, 2)])
def test_validation_batch_sharding_deals_out_the_reference_batches(n, world, bs):
    from dfcsa.data_loader import GpuLoader
    order = list(range(n))
    single = GpuLoader.rank_batches(order, bs, 0, 1, False)
    dealt = [GpuLoader.rank_batches(order, bs, r, world, False, shard_batches=True) for r in range(world)]
    assert sorted(b for br in dealt for b in br) == sorted(single)            # the same batches, each on exactly one rank
    assert all(br == single[r::world] for r, br in enumerate(dealt))


class _FakeVal:
    """Stands in for the model + metric kernels in Trainer.validate_epoch's reduction logic (CPU, gloo)."""
    shard_batches = True

    def __init__(self, batches):
        self.batches = batches

    def __iter__(self):
        return iter(self.batches)


def _val_worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "dfc-sa-unet_b200")]
    from dfcsa import metrics as M
    from dfcsa import trainer as T
    from dfcsa.data_loader import GpuLoader
    from oracle import dfcsa_oracle as O
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)

    def per_sample(p, t, loss_type, params):        # CPU stand-in for the batched CUDA metric kernels
        rows = []
        for j in range(p.shape[0]):
            m = O.calculate_metrics(p[j:j + 1].reshape(1, 1, 1, -1), t[j:j + 1].reshape(1, 1, 1, -1).float(), "bce_dice", {})
            rows.append([float(m["loss"]), 0.0, 0.0, m["iou"], m["dice"]])
        return torch.tensor(rows)
    M.per_sample_metrics = per_sample
    n, bs = 7, 2
    g = torch.Generator().manual_seed(3)
    logit = torch.randn(n, 1, 8, 8, generator=g) * 2
    mask = (torch.rand(n, 1, 8, 8, generator=g) > 0.5).float()
    mine = GpuLoader.rank_batches(list(range(n)), bs, rank, world, False, shard_batches=True)
    batches = [{"image": logit[b], "mask": mask[b], "filename": [f"s{i}" for i in b], "index": b} for b in mine]
    tr = T.Trainer.__new__(T.Trainer)
    tr.world, tr.rank, tr.device = world, rank, torch.device("cpu")
    tr.config = {"logging": {"save_best_worst_samples": 2}}
    tr.loss_type, tr.loss_params = "bce_dice", {}
    tr.model = torch.nn.Identity()                  # "logits" are fed in as the images
    res = tr.validate_epoch(_FakeVal(batches))
    torch.save({k: res[k] for k in ("loss", "iou", "dice")} |
               {"worst": [s["filename"] for s in res["worst_samples"]], "best": [s["filename"] for s in res["best_samples"]]},
               os.path.join(out_dir, f"v{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_gloo_world2_validation_reduces_to_the_full_set_metrics(tmp_path):
    """under data parallel the validation metrics (and with them best_val_loss / is_best) must be those of
    the FULL validation set on every rank, not of rank 0's shard."""
    from oracle import dfcsa_oracle as O
    world = 2
    mp.spawn(_val_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(tmp_path / f"v{r}.pt") for r in range(world)]
    assert outs[0] == outs[1]                                                  # every rank reports the same numbers
    n, bs = 7, 2
    g = torch.Generator().manual_seed(3)
    logit = torch.randn(n, 1, 8, 8, generator=g) * 2
    mask = (torch.rand(n, 1, 8, 8, generator=g) > 0.5).float()
    p = torch.sigmoid(logit)
    per_batch = [O.calculate_metrics(p[i:i + bs], mask[i:i + bs], "bce_dice", {}) for i in range(0, n, bs)]
    for k in ("loss", "iou", "dice"):
        assert abs(outs[0][k] - sum(float(m[k]) for m in per_batch) / len(per_batch)) < 1e-5, k
    dice = sorted((O.calculate_metrics(p[i:i + 1], mask[i:i + 1], "bce_dice", {})["dice"], i) for i in range(n))
    assert outs[0]["worst"] == [f"s{i}" for _, i in dice[:2]] and outs[0]["best"] == [f"s{i}" for _, i in dice[-2:]]
