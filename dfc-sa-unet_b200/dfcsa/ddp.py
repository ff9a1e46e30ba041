"""Data-parallel gradient exchange (new; the reference is single-device - SURVEY.md 8(e)).

One process per GPU, the batch sharded by rank, the model replicated.  All parameter gradients live in ONE flat fp32
buffer in `net.parameters()` order (dfcsa.optim.FusedSGD.flat_grad), so a bucket is a contiguous slice of it and the
exchange is `all_reduce(SUM)` on that slice - no packing copies.  Buckets follow the reverse execution order of the
network: bucket k is reduced on a side stream as soon as net_backward reports that the k-th group of
layers has enqueued its last weight gradient, which hides the decoder and bottleneck buckets (86 % of the bytes) behind
the encoder's backward.  The 1/world factor and the global-norm clip are applied afterwards by the fused SGD kernel.

The reducer is device-agnostic (a CPU flat tensor + the gloo backend work the same way), which is how the host logic
is tested without GPUs (tests/test_ddp_cpu.py).
"""
import torch
import torch.distributed as dist


FLAT_ALIGN = 32      # every tensor starts on a 128-byte boundary of the flat gradient / momentum buffers


def flat_offsets(params, align=FLAT_ALIGN):
    """{param: (lo, hi)} element ranges inside the flat buffers and the padded total length (parameters() order; each
    tensor 128-byte aligned so that the weight-gradient kernels can use 16-byte vector reductions into it)."""
    offs, off = {}, 0
    for p in params:
        offs[p] = (off, off + p.numel())
        off = (off + p.numel() + align - 1) // align * align
    return offs, off


def bucket_modules(net):
    """Groups of sub-modules in the order their backward completes (net_backward's after_stage(k) contract)."""
    return [[net.final_conv, net.up_conv1, net.up1], [net.up_conv2, net.up2], [net.up_conv3, net.up3], [net.up_conv4],
            [net.up4], [net.bottleneck], [net.down4], [net.down3], [net.down2, net.down1]]


def bucket_ranges(net):
    """Per bucket, the maximal contiguous runs [(lo, hi), ...] of its parameters inside the flat gradient buffer
    (parameters() order).  In this network every bucket is a single run (up_k / up_conv_k / final_conv are adjacent
    in registration order); the function still checks that the buckets partition the buffer exactly."""
    offs, total = flat_offsets(list(net.parameters()))
    done = torch.zeros(total, dtype=torch.bool) if total < (1 << 28) else None
    out = []
    for mods in bucket_modules(net):
        segs = sorted(offs[p] for m in mods for p in m.parameters())
        # merge adjacent parameter ranges (alignment padding included) into maximal contiguous runs
        runs = []
        for lo, hi in segs:
            if runs and (runs[-1][1] + FLAT_ALIGN - 1) // FLAT_ALIGN * FLAT_ALIGN == lo:
                runs[-1][1] = hi
            else:
                runs.append([lo, hi])
        out.append([(lo, hi) for lo, hi in runs])
        if done is not None:
            for lo, hi in runs:
                assert not bool(done[lo:hi].any()), "a parameter belongs to two gradient buckets"
                done[lo:hi] = True
    if done is not None:
        for lo, hi in offs.values():
            assert bool(done[lo:hi].all()), "a parameter belongs to no gradient bucket"
    return out


class BucketReducer:
    """all_reduce(SUM) of bucket k of a flat gradient buffer; on CUDA the collective runs on a side stream ordered
    after the compute stream's current position, and finish() joins it back."""

    def __init__(self, flat_grad, ranges, group=None):
        self.flat = flat_grad
        self.ranges = ranges
        self.group = group
        self.cuda = flat_grad.is_cuda
        self.stream = torch.cuda.Stream(device=flat_grad.device) if self.cuda else None
        self.views = [[flat_grad[lo:hi] for lo, hi in runs] for runs in ranges]
        self.bytes = [sum(4 * (hi - lo) for lo, hi in runs) for runs in ranges]

    def reduce(self, k):
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record()
            self.stream.wait_event(ev)
            with torch.cuda.stream(self.stream):
                for v in self.views[k]:
                    dist.all_reduce(v, group=self.group)
        else:
            for v in self.views[k]:
                dist.all_reduce(v, group=self.group)

    def finish(self):
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.stream)
