"""Data-parallel step on real GPUs (needs >= 2 devices; skipped otherwise): two NCCL ranks, each with its own shard,
must end with exactly the parameters a single process gets when it averages the two shards' gradients itself
(per-replica BatchNorm statistics), clips by the global norm and takes the SGD step."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _setup_paths():
    for p in (ROOT, os.path.join(ROOT, "dfc-sa-unet_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)


CFG = {"training": {"loss": {"type": "bce_dice", "params": {}}, "num_epochs": 1}, "logging": {"log_dir": "/tmp/dfcsa_ddp_test"}}


def _model():
    from dfcsa.modules import UNetDFCSARes
    from dfcsa.selftest import set_gamma
    torch.manual_seed(0)
    m = UNetDFCSARes(3, 1, [64, 128, 256, 512], pool_size=4, ablation_on_qk_channels=8)
    set_gamma(m, 0.5)
    return m


def _worker(rank, world, port, out_dir, global_dice=False):
    _setup_paths()
    import copy
    import torch.distributed as dist
    from dfcsa.trainer import Trainer
    from oracle import dfcsa_oracle as O
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    cfg = copy.deepcopy(CFG)
    cfg["training"]["global_batch_dice"] = global_dice
    model = _model()
    if rank == 1:       # a replica that starts from different weights must be overwritten by rank 0's (Trainer.sync_replicas)
        with torch.no_grad():
            model.final_conv.weight.add_(1.0)
    tr = Trainer(model, None, None, None, f"cuda:{rank}", cfg)
    img, mask = O.synthetic_batch(2 * world, 64, 64, seed=7)
    for _ in range(2):
        r = tr.train_step(img[2 * rank:2 * rank + 2].cuda(), mask[2 * rank:2 * rank + 2].cuda())
    torch.cuda.synchronize()
    flat = torch.cat([p.detach().reshape(-1) for p in tr.model.parameters()]).cpu()
    torch.save({"flat": flat, "stats": r.host()}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("global_dice", [False, True])
def test_two_rank_nccl_step_matches_single_process_average(cuda, tmp_path, global_dice):
    """global_dice=True (training.global_batch_dice): the loss sums are all-reduced, so the loss is BCE over the global
    batch + ONE Dice ratio over the global batch; its gradient is sum_r J_r^T [w_bce (p - t) / (world n) + w_dice dDice_global]
    (the Dice part used to come out world times too small)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from dfcsa import engine, ops
    from dfcsa.optim import FusedSGD
    from oracle import dfcsa_oracle as O
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), global_dice), nprocs=world, join=True)
    outs = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    assert torch.equal(outs[0]["flat"], outs[1]["flat"])        # replicas stay bit-identical
    # single-process emulation of the same two steps with the library's own kernels
    net = _model().cuda()
    opt = FusedSGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4, max_norm=1.0)
    img, mask = O.synthetic_batch(2 * world, 64, 64, seed=7)
    for _ in range(2):
        opt.zero_grad()
        acc = torch.zeros_like(opt.flat_grad)
        state = {k: v.clone() for k, v in net.state_dict().items() if "running" in k or "tracked" in k}
        gsums = torch.zeros(8, dtype=torch.float64, device="cuda")
        if global_dice:       # the all-reduced loss sums: a forward pass of every shard first (BN buffers restored below)
            for r in range(world):
                net.load_state_dict(state, strict=False)
                net.train()
                with torch.no_grad():
                    lg, _ = engine.net_forward(net, img[2 * r:2 * r + 2].cuda(), True, save=False)
                ops.bce_dice_sums(lg, mask[2 * r:2 * r + 2].cuda().float(), True, gsums)
        for r in range(world):
            net.load_state_dict(state, strict=False)              # every replica starts from the same BN buffers
            opt.flat_grad.zero_()
            net.train()
            x, t = img[2 * r:2 * r + 2].cuda(), mask[2 * r:2 * r + 2].cuda().float()
            logits, ctx = engine.net_forward(net, x, True, save=True)
            sums = torch.zeros(8, dtype=torch.float64, device="cuda")
            ops.bce_dice_sums(logits, t, True, sums)
            dlogits = torch.empty_like(logits)
            if global_dice:     # exact gradient of the global loss w.r.t. this shard's logits: BCE 1/(world n), Dice from the global sums
                ops.bce_dice_bwd(logits, t, True, gsums, 1.0 / world, 1.0, 1.0, None, dlogits)
            else:
                ops.bce_dice_bwd(logits, t, True, sums, 1.0, 1.0, 1.0, None, dlogits)
            engine.net_backward(net, ctx, dlogits, opt.grads)
            acc += opt.flat_grad
        opt.flat_grad.copy_(acc)
        opt.step(grad_scale=1.0 if global_dice else 1.0 / world)
    ref = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).cpu()
    # atomics in the weight-gradient / BN reductions make the last bits run-to-run variable
    err = (outs[0]["flat"] - ref).norm() / ref.norm()
    assert err < 1e-4, float(err)
