"""Trainer mirror (reference utils/trainer.py:20-460) for the hot loop, plus the data-parallel step.

train_step() is one iteration of Trainer.train_epoch (:115-157): forward, sigmoid + bce_dice, backward,
clip_grad_norm_(1.0), SGD step - all on libdfcsa kernels, with no device->host sync inside (the reference does >= 7
.item() syncs per step; here the five step scalars stay on the device and are read back once per `log_every`).

Data parallel (new; the reference is single-device): one process per GPU, batch sharded by rank, per-replica
BatchNorm statistics (standard DDP semantics), gradients summed with NCCL all-reduce on the flat gradient buffer of
FusedSGD - issued per bucket on a side stream as soon as that bucket's last weight gradient has been enqueued, so the
reduction of the decoder / bottleneck buckets overlaps the encoder's backward - then divided by world size inside
the fused clip + SGD kernel (clip uses the norm of the averaged gradient, exactly as a single-process run would).
"""
import os
import time

import torch
import torch.distributed as dist

from . import engine, ops
from .ddp import BucketReducer, bucket_modules, bucket_ranges
from .metrics import bce_dice_with_logits, calculate_metrics
from .optim import FusedSGD


_DDP_OVERLAP = os.environ.get("DFCSA_DDP_OVERLAP", "1") != "0"     # per-bucket all-reduce overlapped with the backward pass


class StepResult:
    __slots__ = ("stats",)

    def __init__(self, stats):
        self.stats = stats      # device tensor [loss, bce, dice_loss, hard_iou, hard_dice]

    def host(self):
        v = self.stats.tolist()
        return {"loss": v[0], "bce": v[1], "dice_loss": v[2], "iou": v[3], "dice": v[4]}


def reduce_buckets(net):
    """Parameter buckets in reverse execution order: each is reduced as soon as its backward is done."""
    return [[p for m in mods for p in m.parameters()] for mods in bucket_modules(net)]


def device_feeder(batches, device):
    """Yields (images, masks) on the device with ONE batch of look-ahead: the host->device copy of batch i+1 runs on a
    copy stream while batch i trains (the reference copies synchronously inside the loop, utils/trainer.py:116-117).
    `batches` yields (images, masks) host tensors (pinned memory makes the copies asynchronous); batches that already
    live on the device (dfcsa.data_loader) pass straight through.

    The copies land in two persistent staging buffers per shape, re-used alternately: allocating per batch on the copy
    stream makes the caching allocator fall back to cudaMalloc (blocks freed across streams are not immediately
    re-usable), which synchronises the device every step.  A buffer is overwritten only after an event recorded when the
    consumer asks for the NEXT batch, i.e. after everything that reads it has been enqueued."""
    device = torch.device(device)
    copy_stream = torch.cuda.Stream(device=device)
    slots = [None, None]          # [images, masks, consumed event or None]

    def stage(pair, k):
        if pair[0].is_cuda:
            return pair[0], pair[1], None, None
        slot = slots[k]
        if slot is None or slot[0].shape != pair[0].shape or slot[0].dtype != pair[0].dtype or slot[1].shape != pair[1].shape \
                or slot[1].dtype != pair[1].dtype:
            slot = slots[k] = [torch.empty(pair[0].shape, dtype=pair[0].dtype, device=device),
                               torch.empty(pair[1].shape, dtype=pair[1].dtype, device=device), None]
            copy_stream.wait_stream(torch.cuda.current_stream(device))     # the fresh blocks may have had earlier users
        with torch.cuda.stream(copy_stream):
            if slot[2] is not None:
                copy_stream.wait_event(slot[2])
            slot[0].copy_(pair[0], non_blocking=True)
            slot[1].copy_(pair[1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return slot[0], slot[1], ev, slot

    it = iter(batches)
    try:
        nxt = stage(next(it), 0)
    except StopIteration:
        return
    k = 0
    while nxt is not None:
        img, msk, ev, slot = nxt
        cur = torch.cuda.current_stream(device)
        if ev is not None:
            cur.wait_event(ev)
        k ^= 1
        try:
            nxt = stage(next(it), k)
        except StopIteration:
            nxt = None
        yield img, msk
        if slot is not None:                    # the consumer is back for more: its reads of this buffer are enqueued
            slot[2] = torch.cuda.Event()
            slot[2].record(torch.cuda.current_stream(device))


class Trainer:
    def __init__(self, model, train_loader, val_loader, optimizer, device, config, log_every=10):
        self.config = config
        self.device = torch.device(device)
        self.model = model.to(self.device)
        self.train_loader, self.val_loader = train_loader, val_loader
        tr = config["training"]
        self.loss_type = tr.get("loss", {}).get("type", "dice")
        self.loss_params = tr.get("loss", {}).get("params", {})
        if self.loss_type != "bce_dice":
            raise NotImplementedError("dfcsa.Trainer runs the bce_dice loss of the DFC-SA configs")
        # reference quirk kept: calculate_metrics reads weight_bce / weight_dice (utils/metrics.py:246-247)
        self.w_bce = float(self.loss_params.get("weight_bce", 1.0))
        self.w_dice = float(self.loss_params.get("weight_dice", 1.0))
        if optimizer is None or not isinstance(optimizer, FusedSGD):
            base = optimizer.param_groups[0] if optimizer is not None else {}
            optimizer = FusedSGD(self.model.parameters(), lr=float(base.get("lr", tr.get("learning_rate", 0.01))),
                                 momentum=float(base.get("momentum", tr.get("momentum", 0.9))),
                                 weight_decay=float(base.get("weight_decay", tr.get("weight_decay", 1e-4))), max_norm=1.0)
        self.optimizer = optimizer
        self.num_epochs = tr.get("num_epochs", 1)
        self.log_every = log_every
        self.train_losses, self.val_losses = [], []
        self.train_dice_scores, self.val_dice_scores = [], []
        self.train_iou_scores, self.val_iou_scores = [], []
        self.best_val_loss = float("inf")
        log = config.get("logging", {})
        self.log_dir = str(log.get("log_dir", "runs/dfcsa")).replace("\\", "/")
        self.checkpoint_dir = os.path.join(self.log_dir, "checkpoints")
        self.best_model_path = os.path.join(self.log_dir, "best_model.pth")
        self.start_time = time.time()
        # data parallel
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self._reducer = BucketReducer(self.optimizer.flat_grad, bucket_ranges(self.model)) if self.world > 1 else None
        self.sync_replicas()
        self._graphs = {}       # (input shapes) -> [calls, CUDAGraph, static image, static mask, static stats]

    def sync_replicas(self):
        """Data parallel: every replica starts from rank 0's parameters, BatchNorm buffers and momentum (in place, so the
        flat-buffer views and the device pointer tables stay valid).  Identical seeds usually make this a no-op, but a
        rank whose pretrained / resume checkpoint failed to load would otherwise train a diverged replica silently."""
        if self.world <= 1:
            return
        with torch.no_grad():
            for t in list(self.model.parameters()) + list(self.model.buffers()):
                dist.broadcast(t.data, src=0)
            dist.broadcast(self.optimizer.flat_mom, src=0)

    # ------------------------------------------------------------------------------------------------------------
    def train_step(self, images, masks, micro_batches=1):
        """reference utils/trainer.py:116-151 for one batch already on the device.  Returns StepResult (device scalars).

        micro_batches = k > 1 (not in the reference): the batch is run as k consecutive forward/backward passes whose
        gradients are summed before ONE clip + SGD step - for per-GPU batches whose activations exceed one forward's memory
        (C4: 128 / 256 images of 512^2 per GPU).  Loss, BatchNorm statistics and Dice are then those of each micro-batch
        (the gradient is the mean of the k micro-batch gradients, the returned stats their mean)."""
        net, opt = self.model, self.optimizer
        net.train()
        opt.zero_grad()
        k = int(micro_batches)
        if k <= 1:
            overlap = self.world > 1 and self.config.get("training", {}).get("ddp_overlap", _DDP_OVERLAP)
            stats = self._forward_backward(images, masks, self._reducer.reduce if overlap else None)
            if self.world > 1 and not overlap:      # every bucket after the backward pass, nothing shares the SMs with it
                for b in range(len(self._reducer.ranges)):
                    self._reducer.reduce(b)
        else:
            if images.shape[0] % k:
                raise ValueError(f"dfcsa: batch of {images.shape[0]} does not split into {k} equal micro-batches")
            if getattr(self, "_accum", None) is None:
                self._accum = torch.empty_like(opt.flat_grad)
            stats = None
            for i, (im, mk) in enumerate(zip(images.chunk(k), masks.chunk(k))):
                st = self._forward_backward(im, mk, None)
                stats = st if stats is None else stats + st
                if i == 0:
                    self._accum.copy_(opt.flat_grad)
                elif i < k - 1:
                    ops.accumulate(self._accum, opt.flat_grad)
                if i < k - 1:
                    opt.flat_grad.zero_()           # the kernels expect zero-initialised gradient tensors
            ops.accumulate(opt.flat_grad, self._accum)
            stats = stats / k
            if self.world > 1:                      # all buckets after the last micro-batch (1 ms of NVLink time per step)
                for b in range(len(self._reducer.ranges)):
                    self._reducer.reduce(b)
        if self.world > 1:
            self._reducer.finish()
        opt.step(grad_scale=1.0 / (self.world * max(k, 1)))
        return StepResult(stats)

    def _forward_backward(self, images, masks, after_stage):
        """forward, sigmoid + bce_dice, backward of one (micro-)batch into the optimizer's flat gradient buffer."""
        net, opt = self.model, self.optimizer
        logits, ctx = engine.net_forward(net, images, True, save=True)
        dev = logits.device
        n = logits.numel()
        sums = torch.zeros(8, dtype=torch.float64, device=dev)
        stats = torch.empty(5, dtype=torch.float32, device=dev)
        t = masks.contiguous().float()
        ops.bce_dice_sums(logits, t, True, sums)
        w_dice_bwd = self.w_dice
        if self.world > 1 and self.config.get("training", {}).get("global_batch_dice", False):
            dist.all_reduce(sums)           # exact global-batch Dice / BCE; n scales with world
            n = n * self.world
            # With global sums bce_dice_bwd yields the exact derivative of the GLOBAL Dice term for every local logit, so
            # the NCCL SUM over ranks is already the whole Dice gradient; the BCE term (local 1/n) still needs the
            # 1/world average.  step(grad_scale=1/world) divides both: pre-multiply the Dice term by world.
            w_dice_bwd = self.w_dice * self.world
        ops.bce_dice_finalize(sums, n, self.w_bce, self.w_dice, 1.0, stats)
        dlogits = torch.empty_like(logits)
        ops.bce_dice_bwd(logits, t, True, sums, self.w_bce, w_dice_bwd, 1.0, None, dlogits)
        engine.net_backward(net, ctx, dlogits, opt.grads, after_stage=after_stage)
        return stats

    def train_step_graphed(self, images, masks, micro_batches=1):
        """train_step replayed from a CUDA graph (one graph per input shape).  Every libdfcsa entry point only enqueues
        work and all buffers come from PyTorch's allocator, so the ~600 launches of a step are captured as they are;
        replaying them removes the host-side launch cost that dominates small batches (the reference's batch-4 config).
        The first two calls per shape run eagerly (lazy one-time state settles: packed-weight buffers, momentum
        initialisation, kernel attributes); the third call captures and replays.  images / masks may live in (pinned)
        host memory: they are copied straight into the graph's static input buffers."""
        key = (tuple(images.shape), images.dtype, tuple(masks.shape), masks.dtype, int(micro_batches))
        from . import _lib
        g = self._graphs.setdefault(key, [0, None, None, None, None, 0])
        g[0] += 1
        if g[0] <= 2:
            return self.train_step(images.to(self.device, non_blocking=True), masks.to(self.device, non_blocking=True), micro_batches)
        if g[1] is None:
            g[2], g[3] = images.to(self.device).clone(), masks.to(self.device).clone()
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            n0 = _lib.LAUNCHES
            with torch.cuda.graph(graph):
                g[4] = self.train_step(g[2], g[3], micro_batches).stats
            g[5] = _lib.LAUNCHES - n0        # libdfcsa kernels inside the graph (capturing does not run them)
            _lib.LAUNCHES = n0
            g[1] = graph
        g[2].copy_(images, non_blocking=True)
        g[3].copy_(masks, non_blocking=True)
        g[1].replay()
        _lib.LAUNCHES += g[5]
        return StepResult(g[4])

    # ------------------------------------------------------------------------------------------------------------
    def train_epoch(self, epoch):
        running = torch.zeros(5, dtype=torch.float32, device=self.device)
        good = torch.zeros((), dtype=torch.float32, device=self.device)
        nb = 0
        # training.cuda_graph: replay the step from a CUDA graph (one per batch shape; a ragged last batch is captured
        # separately) - worth it when the batch is small enough for launch overhead to show
        step = self.train_step_graphed if self.config.get("training", {}).get("cuda_graph", False) else self.train_step
        for images, masks in device_feeder(((b["image"], b["mask"]) for b in self.train_loader), self.device):
            r = step(images, masks)
            # a batch with a non-finite loss is skipped by the device-side update (sgd_step) AND left out of the epoch
            # means, like the reference's `continue` before accumulating (utils/trainer.py:134-139); no host sync
            ok = torch.isfinite(r.stats[0])
            running += torch.where(ok, r.stats, torch.zeros_like(r.stats))
            good += ok.to(torch.float32)
            nb += 1
        v = (running / good.clamp(min=1.0)).tolist()
        self.skipped_batches = nb - int(good.item())
        return v[0], v[3], v[4]

    @torch.no_grad()
    def validate_epoch(self, dataloader):
        """reference utils/trainer.py:172-265: mean loss / IoU / Dice over the batches plus the K best and K worst samples
        by Dice (K = logging.save_best_worst_samples).  The reference re-evaluates every sample separately and copies
        image, mask and prediction of EVERY sample to the host each epoch; here the per-sample metrics are one batched
        kernel pair, the candidates for best / worst stay in two device-side pools of at most 2K samples, nothing is
        synchronised inside the loop, and only the 2K winners are copied out at the end."""
        from .metrics import per_sample_metrics
        self.model.eval()
        K = int(self.config.get("logging", {}).get("save_best_worst_samples", 0) or 0)
        dev = self.device
        tot = torch.zeros(3, dtype=torch.float64, device=dev)
        nb, seen, names, gidx = 0, 0, [], []
        pool = None                 # (ids, metrics [m, 5], images, masks, outputs)
        for batch in dataloader:
            images = batch["image"].to(dev, non_blocking=True)
            masks = batch["mask"].to(dev, non_blocking=True)
            outputs = torch.sigmoid(self.model(images))
            sm = per_sample_metrics(outputs, masks, self.loss_type, self.loss_params)
            bm = per_sample_metrics(outputs.reshape(1, -1), masks.reshape(1, -1), self.loss_type, self.loss_params)[0]
            ok = ~torch.isnan(bm[0])                                   # a NaN batch is skipped (reference :210-212)
            tot += torch.where(ok, bm[[0, 3, 4]].double(), torch.zeros(3, dtype=torch.float64, device=dev))
            nb += 1
            b = images.shape[0]
            names += list(batch.get("filename", [str(seen + i) for i in range(b)]))
            gidx += [int(i) for i in batch.get("index", range(seen, seen + b))]      # position in the full validation set
            if K > 0:
                ids = torch.arange(seen, seen + b, device=dev)
                sm = torch.where(ok, sm, torch.full_like(sm, float("nan")))
                cand = (ids, sm, images, masks.float(), outputs)
                pool = cand if pool is None else tuple(torch.cat([p, c]) for p, c in zip(pool, cand))
                if pool[0].numel() > 2 * K:
                    d = pool[1][:, 4]
                    worst = torch.topk(torch.nan_to_num(d, nan=float("inf")), K, largest=False).indices
                    best = torch.topk(torch.nan_to_num(d, nan=float("-inf")), K, largest=True).indices
                    keep = torch.unique(torch.cat([worst, best]))
                    pool = tuple(p[keep] for p in pool)
            seen += b
        # Data parallel with a batch-sharded loader (GpuLoader(shard_batches=True)): every rank holds the per-batch sums
        # of its share of the reference's batches; one all-reduce of (sums, batch count) gives every rank the
        # full-set means the reference computes, so best_val_loss / is_best / the histories agree across ranks.  The
        # collectives sit after the loop, so ranks with different batch counts cannot dead-lock.
        sharded = self.world > 1 and getattr(dataloader, "shard_batches", False)
        if sharded:
            red = torch.cat([tot, torch.tensor([float(nb)], dtype=torch.float64, device=dev)])
            dist.all_reduce(red)
            tot, nb = red[:3], int(round(float(red[3])))
        mean = (tot / max(nb, 1)).tolist()                             # the epoch's only synchronisation
        res = {"loss": mean[0], "iou": mean[1], "dice": mean[2], "best_samples": [], "worst_samples": []}
        if K > 0 and (pool is not None or sharded):
            cands = []
            if pool is not None:
                ids, sm, imgs, msks, outs = (p.cpu() for p in pool)
                valid = ~torch.isnan(sm[:, 4])
                order = [i for i in torch.argsort(sm[:, 4], stable=True).tolist() if valid[i]]      # ascending Dice (:251)

                def sample(i):
                    return {"sample_idx": gidx[int(ids[i])], "image": imgs[i], "mask": msks[i], "output": outs[i],
                            "filename": names[int(ids[i])],
                            "metrics": {"loss": float(sm[i, 0]), "iou": float(sm[i, 3]), "dice": float(sm[i, 4])}}
                cands = [sample(i) for i in order[:K]] + [sample(i) for i in order[K:][-K:]]
            if sharded:      # K worst + K best of every rank -> the global K worst / K best (ties: dataset order, as a stable sort)
                allc = [None] * self.world
                dist.all_gather_object(allc, cands)
                cands = [c for rank_c in allc for c in rank_c]
            cands.sort(key=lambda c: (c["metrics"]["dice"], c["sample_idx"]))
            res["worst_samples"] = cands[:K]
            res["best_samples"] = cands[-K:]
        return res

    def save_checkpoint(self, epoch, metrics, is_best=False):
        """reference utils/trainer.py:267-298 (same dict keys / file names); rank 0 only under data parallel."""
        if self.rank != 0:
            return
        os.makedirs(self.checkpoint_dir, exist_ok=True)
        ckpt = {"epoch": epoch, "model_state_dict": self.model.state_dict(), "optimizer_state_dict": self.optimizer.state_dict(),
                "train_losses": self.train_losses, "val_losses": self.val_losses,
                "train_dice_scores": self.train_dice_scores, "val_dice_scores": self.val_dice_scores,
                "train_iou_scores": self.train_iou_scores, "val_iou_scores": self.val_iou_scores,
                "best_val_loss": self.best_val_loss, "metrics": metrics}
        torch.save(ckpt, os.path.join(self.checkpoint_dir, f"checkpoint_epoch_{epoch + 1}.pth"))
        if is_best:
            torch.save(self.model.state_dict(), self.best_model_path)
            torch.save(ckpt, os.path.join(self.checkpoint_dir, "best_checkpoint.pth"))

    def load_checkpoint(self, checkpoint_path):
        """reference utils/trainer.py:300-324."""
        ckpt = torch.load(checkpoint_path.replace("\\", "/"), map_location=self.device, weights_only=False)
        self.model.load_state_dict(ckpt["model_state_dict"])
        self.optimizer.load_state_dict(ckpt["optimizer_state_dict"])
        for k in ("train_losses", "val_losses", "train_dice_scores", "val_dice_scores", "train_iou_scores", "val_iou_scores",
                  "best_val_loss"):
            setattr(self, k, ckpt[k])
        self.sync_replicas()
        return ckpt["epoch"]

    def train(self, resume_from=None):
        """reference utils/trainer.py:326-460.  resume_from: checkpoint path; training continues at the epoch after the
        checkpoint's with the metric histories it stored (the reference indexes the integer load_checkpoint returns,
        :337-339, and then clears the histories, :342-350, so its resume path cannot run - this is the evident intent)."""
        start_epoch = 0
        if resume_from:
            start_epoch = self.load_checkpoint(resume_from) + 1
            if self.rank == 0:
                print(f"從 epoch {start_epoch} 恢復訓練")
        self.epochs = list(range(1, len(self.train_losses) + 1))
        best = max(self.val_dice_scores) if self.val_dice_scores else 0.0
        for epoch in range(start_epoch, self.num_epochs):
            tl, ti, td = self.train_epoch(epoch)
            self.epochs.append(epoch + 1)
            self.train_losses.append(tl); self.train_iou_scores.append(ti); self.train_dice_scores.append(td)
            if self.rank == 0:
                print(f"Epoch [{epoch + 1}/{self.num_epochs}]")
                print(f"  Train Loss: {tl:.4f}, Dice: {td:.4f}, IoU: {ti:.4f}")
            if self.val_loader is None:
                continue
            val = self.validate_epoch(self.val_loader)
            self.val_losses.append(val["loss"]); self.val_iou_scores.append(val["iou"]); self.val_dice_scores.append(val["dice"])
            self.best_val_loss = min(self.best_val_loss, val["loss"])
            is_best = val["dice"] > best
            best = max(best, val["dice"])
            if self.rank == 0:
                print(f"  Val Loss: {val['loss']:.4f}, Dice: {val['dice']:.4f}, IoU: {val['iou']:.4f}")
            freq = self.config["training"].get("save_checkpoint_freq", 100)
            if (epoch + 1) % freq == 0 or is_best:
                self.save_checkpoint(epoch, _slim_metrics(val), is_best)
            if self.rank == 0 and (val["best_samples"] or val["worst_samples"]):
                for kind in ("best_samples", "worst_samples"):
                    save_prediction_samples(val[kind], os.path.join(self.log_dir, f"epoch_{epoch + 1}", kind))
        if self.rank == 0:
            total = time.time() - self.start_time
            print(f"Training completed in {int(total // 3600)}h {int(total % 3600 // 60)}m {int(total % 60)}s")
            print(f"Best validation dice: {best:.4f}")


def _slim_metrics(val):
    """What goes into a checkpoint's 'metrics' entry: the numbers and file names, not the image tensors the reference
    pickles into every checkpoint (utils/trainer.py:287 stores validate_epoch's whole return value)."""
    slim = {k: val[k] for k in ("loss", "iou", "dice")}
    for kind in ("best_samples", "worst_samples"):
        slim[kind] = [{"filename": s["filename"], "metrics": s["metrics"]} for s in val.get(kind, [])]
    return slim


def save_prediction_samples(samples, out_dir, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """One PNG per sample: de-normalised image | ground truth | prediction (reference utils/visualization.py via
    utils/trainer.py:413-447).  Needs Pillow; silently skipped without it (logging only, not on the hot path)."""
    try:
        from PIL import Image
    except ImportError:
        return
    os.makedirs(out_dir, exist_ok=True)
    m, s = torch.tensor(mean).view(3, 1, 1), torch.tensor(std).view(3, 1, 1)
    for smp in samples:
        img = smp["image"].float()
        img = (img[:3] * s + m).clamp(0, 1) if img.shape[0] >= 3 else img[:1].expand(3, -1, -1).clamp(0, 1)
        panels = [img, smp["mask"].float().expand(3, -1, -1), smp["output"].float().expand(3, -1, -1)]
        strip = (torch.cat(panels, dim=2) * 255).round().byte().permute(1, 2, 0).contiguous().numpy()
        stem = os.path.splitext(os.path.basename(str(smp["filename"])))[0]
        Image.fromarray(strip).save(os.path.join(out_dir, f"{stem}_dice{smp['metrics']['dice']:.3f}.png"))
