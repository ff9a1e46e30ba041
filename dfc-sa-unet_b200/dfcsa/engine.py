"""Host-side orchestration of the DFC-SA hot path: which libdfcsa kernel runs on which buffer, in which order.

Mirrors, kernel launch by kernel launch, the reference forward (models/unet_dfc_sa_res.py:95-116 block, :161-204 net)
and the hand-derived backward of the same graph.  PyTorch is used only for device memory
(torch.empty / zeros), streams and a few scalar copies; all arithmetic is in the CUDA library.

Layout: activations NHWC fp16 [M, C] views (M = B*H*W), gradients NHWC bf16.  kind::f16 tcgen05.mma cannot mix fp16
with bf16 operands (measured, tools/tc_probe.cu probe 7); the weight-gradient kernel therefore converts its fp16
activation tiles to bf16 in shared memory, so no bf16 copy of any activation is ever written to HBM.  torch.cat of the
reference is zero-copy: producers write channel slices of one buffer ([f | L | A] inside a block, [up | skip] in the
decoder).
"""
import os as _os

import torch

from . import ops
from .ops import BACKEND_SIMT, BACKEND_TC, OUT_CONVT2x2, TAP_1x1, TAP_2x2S2, TAP_3x3

F16, BF16, F32, F64 = torch.float16, torch.bfloat16, torch.float32, torch.float64
BN_MOMENTUM, BN_EPS = 0.1, 1e-5


def _e(shape, dtype, dev):
    return torch.empty(shape, dtype=dtype, device=dev)


class ZeroArena:
    """Zero-initialised scratch of one pass (forward or backward) of a network, carved from ONE persistent buffer that a
    single memset clears at the start of the pass: the BatchNorm statistic sums, the backward reduction buffers and the
    packed 3x3 / ConvT weight-gradient accumulators were ~70 separate torch.zeros fill kernels per step (250 launches
    in the round-1 smoke trace).  The first pass sizes the buffer (requests fall back to torch.zeros until it exists)."""
    ALIGN = 256

    def __init__(self, dev):
        self.dev, self.buf, self.off, self.need = dev, None, 0, 0

    def begin(self):
        if self.need > (self.buf.numel() if self.buf is not None else 0):
            self.buf = torch.zeros(self.need, dtype=torch.uint8, device=self.dev)
        elif self.buf is not None and self.off > 0:
            self.buf[:self.off].zero_()
        self.off = self.need = 0

    def take(self, shape, dtype):
        n = 1
        for d in shape:
            n *= int(d)
        nbytes = (n * torch.empty((), dtype=dtype).element_size() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.need += nbytes
        if self.buf is None or self.off + nbytes > self.buf.numel():
            return torch.zeros(shape, dtype=dtype, device=self.dev)
        v = self.buf[self.off:self.off + n * torch.empty((), dtype=dtype).element_size()].view(dtype).view(shape)
        self.off += nbytes
        return v


_ARENA = None        # the ZeroArena of the pass being enqueued (net_forward / net_backward set it)


def _z(shape, dtype, dev):
    if _ARENA is not None and _ARENA.dev == dev:
        return _ARENA.take(shape if isinstance(shape, (tuple, list)) else (shape,), dtype)
    return torch.zeros(shape, dtype=dtype, device=dev)


def _net_arena(net, which, dev):
    a = net.__dict__.setdefault("_dfcsa_arena", {}).get(which)
    if a is None or a.dev != dev:
        a = net.__dict__["_dfcsa_arena"][which] = ZeroArena(dev)
    return a


def _backend(segs, w, N, out):
    if ops.tc_eligible(segs, N, out) and w.dtype == segs[0][0].dtype:
        return BACKEND_TC
    return BACKEND_SIMT


class BlockParams:
    """Views of one DynamicFusionConvAttnBlock's parameters / buffers (reference names)."""

    def __init__(self, mod):
        cb, ab, g, fu = mod.conv_branch, mod.attn_branch, mod.gate, mod.fusion_conv
        self.mod = mod
        self.W1, self.b1, self.bn1 = cb[0].weight, cb[0].bias, cb[1]
        self.W2, self.b2, self.bn2 = ab[0].weight, ab[0].bias, ab[1]
        att = ab[3]
        self.att = att
        self.gamma = att.gamma
        self.Wq, self.bq = att.query_conv.weight, att.query_conv.bias
        self.Wk, self.bk = att.key_conv.weight, att.key_conv.bias
        self.Wv, self.bv = att.value_conv.weight, att.value_conv.bias
        self.W3, self.b3, self.bn3 = g[0].weight, g[0].bias, g[1]
        self.W4, self.b4, self.bn4 = fu[0].weight, fu[0].bias, fu[1]
        # in_channels == out_channels: the reference uses nn.Identity (models/unet_dfc_sa_res.py:87-90); here the residual
        # GEMM then multiplies by an identity matrix (never the case in the DFC-SA-Res-Block network itself)
        self.W5 = getattr(mod.residual_conv, "weight", None)
        self.res_scale = mod.res_scale
        self.Ci, self.C = self.W1.shape[1], self.W1.shape[0]
        self.Cq = self.Wq.shape[0]
        self.P = att.pool_size
        self.tc = (self.Ci % 64 == 0) and (self.C % 64 == 0)


    kind = "dfc"

    def bns(self):
        return [self.bn1, self.bn2, self.bn3, self.bn4]


class LocalBlockParams:
    """LocalOnlyBlock (reference models/unet_dfc_sa_ablation_branches.py:72-101): relu(bn(conv3x3(x))) + res_scale * res(x),
    the standard U-Net convolution block of the placement / branch ablations."""

    kind = "local"

    def __init__(self, mod):
        cb = mod.conv_branch
        self.mod = mod
        self.W1, self.b1, self.bn1 = cb[0].weight, cb[0].bias, cb[1]
        self.W5 = getattr(mod.residual_conv, "weight", None)
        self.res_scale = mod.res_scale
        self.Ci, self.C = self.W1.shape[1], self.W1.shape[0]
        self.tc = (self.Ci % 64 == 0) and (self.C % 64 == 0)

    def bns(self):
        return [self.bn1]


class BranchBlockParams:
    """The ablation blocks that keep the attention branch but not the gate:
      "attn"    AttentionOnlyBlock   (reference models/unet_dfc_sa_ablation_branches.py:42-69)   y = A + rs*res
      "add"     AdditionFusionBlock  (models/unet_dfc_sa_ablation_fusion.py:9-55)               y = L + A + rs*res
      "concat"  ConcatFusionBlock    (models/unet_dfc_sa_ablation_fusion.py:58-108)             y = relu(bn(W4 [L|A])) + rs*res
    with L = relu(bn1(conv3x3 x)), A = gamma * up(attn(pool(a))) + a, a = relu(bn2(conv1x1 x))."""

    def __init__(self, mod, kind):
        self.mod, self.kind = mod, kind
        ab = mod.attn_branch
        self.W2, self.b2, self.bn2 = ab[0].weight, ab[0].bias, ab[1]
        att = ab[3]
        self.att, self.gamma = att, att.gamma
        self.Wq, self.bq = att.query_conv.weight, att.query_conv.bias
        self.Wk, self.bk = att.key_conv.weight, att.key_conv.bias
        self.Wv, self.bv = att.value_conv.weight, att.value_conv.bias
        self.has_l = kind != "attn"
        if self.has_l:
            cb = mod.conv_branch
            self.W1, self.b1, self.bn1 = cb[0].weight, cb[0].bias, cb[1]
        if kind == "concat":
            fu = mod.fusion_conv
            self.W4, self.b4, self.bn4 = fu[0].weight, fu[0].bias, fu[1]
        self.W5 = getattr(mod.residual_conv, "weight", None)
        self.res_scale = mod.res_scale
        self.Ci, self.C = self.W2.shape[1], self.W2.shape[0]
        self.Cq = self.Wq.shape[0]
        self.P = att.pool_size
        self.tc = (self.Ci % 64 == 0) and (self.C % 64 == 0)

    def bns(self):
        return ([self.bn1] if self.has_l else []) + [self.bn2] + ([self.bn4] if self.kind == "concat" else [])


def make_block_params(mod):
    """The engine-side view of one block module, by its structure (sub-module names are the reference's)."""
    if hasattr(mod, "gate"):
        return BlockParams(mod)
    if hasattr(mod, "attn_branch"):
        if hasattr(mod, "fusion_conv"):
            return BranchBlockParams(mod, "concat")
        return BranchBlockParams(mod, "add" if hasattr(mod, "conv_branch") else "attn")
    if hasattr(mod, "conv_branch") and not hasattr(mod, "attn_branch"):
        return LocalBlockParams(mod)
    raise NotImplementedError(f"dfcsa: no kernels for block type {type(mod).__name__}")


class PackPlan:
    """Every weight re-layout of a network (fp32 masters -> packed 16-bit GEMM operands) as a device table of jobs that
    ONE dfcsa_pack_jobs launch executes per step.  Destination buffers are persistent, so the plan is built once per
    (network, mode) and only re-run after the optimizer has changed the masters."""

    def __init__(self, dev):
        self.dev = dev
        self.jobs = []
        self.keep = []          # tensors whose storage the table points into
        self._table = None

    def add(self, src, dst, D, s, flip1=False, scale=None, ld_dst=0, row_scale=None):
        self.jobs.append((src, dst, tuple(D), tuple(s), flip1, scale, ld_dst or D[1] * D[2], row_scale))
        self.keep += [src, dst, scale, row_scale]

    def run(self):
        if not self.jobs:
            return
        if self._table is None:
            from . import _lib as L
            tab = (L.PackJob * len(self.jobs))()
            prefix, tot = [0], 0
            for i, (src, dst, D, st, flip1, scale, ld, row_scale) in enumerate(self.jobs):
                j = tab[i]
                j.src, j.dst = src.data_ptr(), dst.data_ptr()
                j.scale = scale.data_ptr() if scale is not None else None
                j.row_scale = row_scale.data_ptr() if row_scale is not None else None
                j.D0, j.D1, j.D2 = D
                j.s0, j.s1, j.s2 = st
                j.ld_dst = ld
                j.src_dtype, j.dst_dtype, j.flip1 = L.dt(src), L.dt(dst), 1 if flip1 else 0
                tot += (D[0] * D[1] * D[2] + 1023) // 1024
                prefix.append(tot)
            self._table = torch.frombuffer(bytearray(bytes(tab)), dtype=torch.uint8).to(self.dev)
            self._prefix = torch.tensor(prefix, dtype=torch.int64, device=self.dev)
            self._chunks = tot
        ops.pack_jobs(self._table, len(self.jobs), self._prefix, self._chunks)


def pack_block_weights(bp, training, need_dx, plan):
    """fp32 master weights -> packed K-major GEMM operands (forward: fp16 / fp32; dgrad: bf16 / fp32): allocates the
    persistent destination buffers and records the re-layout jobs in `plan`."""
    dev = bp.res_scale.device
    Ci, C = bp.Ci, bp.C
    fdt = F16 if bp.tc else F32
    pk = {}
    W5 = bp.W5.detach() if bp.W5 is not None else torch.eye(C, dtype=F32, device=dev).view(C, C, 1, 1)
    if bp.kind in ("attn", "add", "concat"):
        return _pack_branch_block(bp, pk, W5, training, need_dx, plan)
    pk["w1"] = _e((C, 9 * Ci), fdt, dev)
    plan.add(bp.W1.detach(), pk["w1"], (C, 9, Ci), (Ci * 9, 1, 9))
    if bp.kind == "local":
        pk["w5"] = _e((C, Ci), fdt, dev)
        plan.add(W5, pk["w5"], (C, 1, Ci), (Ci, 0, 1))
        if training and need_dx:
            wd = _e((Ci, 10 * C), BF16 if bp.tc else F32, dev)       # [ci, (flipped tap, co) | co (res_scale*W5)]
            plan.add(bp.W1.detach(), wd, (Ci, 9, C), (9, 1, Ci * 9), flip1=True, ld_dst=10 * C)
            plan.add(W5, wd[:, 9 * C:], (Ci, 1, C), (1, 0, Ci), scale=bp.res_scale.detach().reshape(1), ld_dst=10 * C)
            pk["wd15"] = wd
        return pk
    pk["w25"] = _e((2 * C, Ci), fdt, dev)
    plan.add(bp.W2.detach(), pk["w25"][:C], (C, 1, Ci), (Ci, 0, 1))
    plan.add(W5, pk["w25"][C:], (C, 1, Ci), (Ci, 0, 1))
    gdt = F16 if C % 64 == 0 else F32
    pk["w3"] = _e((C, 2 * C), gdt, dev)
    plan.add(bp.W3.detach(), pk["w3"], (C, 1, 2 * C), (2 * C, 0, 1))
    pk["w4"] = _e((C, 3 * C), gdt, dev)
    plan.add(bp.W4.detach(), pk["w4"], (C, 1, 3 * C), (3 * C, 0, 1))
    _pack_attention(bp, pk, training, plan)
    if training:
        bdt = BF16 if C % 64 == 0 else F32
        W4m, W3m = bp.W4.detach().view(C, 3 * C), bp.W3.detach().view(C, 2 * C)
        pk["wd4f"] = _e((C, C), bdt, dev)           # df = dF0 . W4^T[0:C, :]:  wd4f[k, co] = W4[co, k]
        plan.add(W4m, pk["wd4f"], (C, 1, C), (1, 0, 3 * C))
        # [dL | dA] = [dF0 | dG0] . wd43^T with wd43[n, (co of W4 | co of W3)] = (W4[co, C + n] | W3[co, n])
        pk["wd43"] = _e((2 * C, 2 * C), bdt, dev)
        plan.add(W4m[:, C:], pk["wd43"], (2 * C, 1, C), (1, 0, 3 * C), ld_dst=2 * C)
        plan.add(W3m, pk["wd43"][:, C:], (2 * C, 1, C), (1, 0, 2 * C), ld_dst=2 * C)
        if need_dx:   # the first block never needs the gradient w.r.t. the image
            pk["wd125"] = _pack_input_dgrad(bp, W5, plan, True)
    return pk


WEIGHTS_EPOCH = 0    # bumped whenever a libdfcsa kernel rewrites parameters / BatchNorm buffers behind torch's back (raw pointers)


def pack_block_weights_folded(bp, plan, affines):
    """Inference-mode operands of one DFC-SA block (reference inference.py:100 runs the network in eval mode): every
    BatchNorm uses its running statistics, so it is an affine map per channel, y = s * conv(x) + t with
    s = gamma / sqrt(running_var + eps), t = beta + (conv_bias - running_mean) * s.  s is folded into the rows of the packed
    weight matrix (dfcsa_pack_jobs row_scale), t is the GEMM's bias, ReLU is the GEMM's epilogue: conv + BN + ReLU is ONE
    launch and the four affine / activation passes of the training path disappear.
    `affines` collects (bn, conv_bias, scale_out, shift_out) for dfcsa_bn_eval_affine, run before the pack jobs."""
    dev = bp.res_scale.device
    Ci, C = bp.Ci, bp.C
    fdt = F16 if bp.tc else F32
    gdt = F16 if C % 64 == 0 else F32
    pk = {"folded": True}
    W5 = bp.W5.detach() if bp.W5 is not None else torch.eye(C, dtype=F32, device=dev).view(C, C, 1, 1)
    s = _e((4, C), F32, dev)
    pk["t1"], pk["t3"], pk["t4"] = _e((C,), F32, dev), _e((C,), F32, dev), _e((C,), F32, dev)
    pk["b25"] = torch.zeros((2 * C,), dtype=F32, device=dev)          # [t2 | 0]: the residual conv has no bias and no BN
    affines += [(bp.bn1, bp.b1, s[0], pk["t1"]), (bp.bn2, bp.b2, s[1], pk["b25"][:C]), (bp.bn3, bp.b3, s[2], pk["t3"]),
                (bp.bn4, bp.b4, s[3], pk["t4"])]
    pk["w1"] = _e((C, 9 * Ci), fdt, dev)
    plan.add(bp.W1.detach(), pk["w1"], (C, 9, Ci), (Ci * 9, 1, 9), row_scale=s[0])
    pk["w25"] = _e((2 * C, Ci), fdt, dev)
    plan.add(bp.W2.detach(), pk["w25"][:C], (C, 1, Ci), (Ci, 0, 1), row_scale=s[1])
    plan.add(W5, pk["w25"][C:], (C, 1, Ci), (Ci, 0, 1))
    pk["w3"] = _e((C, 2 * C), gdt, dev)
    plan.add(bp.W3.detach(), pk["w3"], (C, 1, 2 * C), (2 * C, 0, 1), row_scale=s[2])
    pk["w4"] = _e((C, 3 * C), gdt, dev)
    plan.add(bp.W4.detach(), pk["w4"], (C, 1, 3 * C), (3 * C, 0, 1), row_scale=s[3])
    _pack_attention(bp, pk, False, plan)
    pk["ones"], pk["zeros"] = torch.ones((C,), dtype=F32, device=dev), torch.zeros((C,), dtype=F32, device=dev)
    pk["scales"] = s
    return pk


def _pack_input_dgrad(bp, W5, plan, has_l):
    """dgrad operand of the convs that read the block input: [ci, (flipped tap, co of W1) | co (W2) | co (res_scale*W5)]."""
    Ci, C = bp.Ci, bp.C
    n = (11 if has_l else 2) * C
    wd = _e((Ci, n), BF16 if bp.tc else F32, W5.device)
    o = 0
    if has_l:
        plan.add(bp.W1.detach(), wd, (Ci, 9, C), (9, 1, Ci * 9), flip1=True, ld_dst=n)
        o = 9 * C
    plan.add(bp.W2.detach(), wd[:, o:], (Ci, 1, C), (1, 0, Ci), ld_dst=n)
    plan.add(W5, wd[:, o + C:], (Ci, 1, C), (1, 0, Ci), scale=bp.res_scale.detach().reshape(1), ld_dst=n)
    return wd


def _pack_branch_block(bp, pk, W5, training, need_dx, plan):
    dev = W5.device
    Ci, C = bp.Ci, bp.C
    fdt = F16 if bp.tc else F32
    gdt = F16 if C % 64 == 0 else F32
    if bp.has_l:
        pk["w1"] = _e((C, 9 * Ci), fdt, dev)
        plan.add(bp.W1.detach(), pk["w1"], (C, 9, Ci), (Ci * 9, 1, 9))
    pk["w25"] = _e((2 * C, Ci), fdt, dev)
    plan.add(bp.W2.detach(), pk["w25"][:C], (C, 1, Ci), (Ci, 0, 1))
    plan.add(W5, pk["w25"][C:], (C, 1, Ci), (Ci, 0, 1))
    if bp.kind == "concat":
        pk["w4c"] = _e((C, 2 * C), gdt, dev)
        plan.add(bp.W4.detach(), pk["w4c"], (C, 1, 2 * C), (2 * C, 0, 1))
    _pack_attention(bp, pk, training, plan)
    if training:
        if bp.kind == "concat":        # [dL | dA] = dF0 . W4:  wd4c[n, co] = W4[co, n]
            pk["wd4c"] = _e((2 * C, C), BF16 if C % 64 == 0 else F32, dev)
            plan.add(bp.W4.detach().view(C, 2 * C), pk["wd4c"], (2 * C, 1, C), (1, 0, 2 * C))
        if need_dx:
            pk["wd125"] = _pack_input_dgrad(bp, W5, plan, bp.has_l)
    return pk


def _pack_attention(bp, pk, training, plan):
    """q/k/v 1x1 convs on the pooled map as ONE GEMM: rows [Wq ; Wk ; Wv], bias [bq | bk | bv] (+ the dgrad operand)."""
    dev = bp.Wq.device
    C = bp.C
    gdt = F16 if C % 64 == 0 else F32
    Cq = bp.Cq
    nq = 2 * Cq + C
    pk["wqkv"] = _e((nq, C), gdt, dev)
    plan.add(bp.Wq.detach(), pk["wqkv"][:Cq], (Cq, 1, C), (C, 0, 1))
    plan.add(bp.Wk.detach(), pk["wqkv"][Cq:2 * Cq], (Cq, 1, C), (C, 0, 1))
    plan.add(bp.Wv.detach(), pk["wqkv"][2 * Cq:], (C, 1, C), (C, 0, 1))
    pk["bqkv"] = _e((nq,), F32, dev)
    plan.add(bp.bq.detach(), pk["bqkv"][:Cq], (1, 1, Cq), (0, 0, 1))
    plan.add(bp.bk.detach(), pk["bqkv"][Cq:2 * Cq], (1, 1, Cq), (0, 0, 1))
    plan.add(bp.bv.detach(), pk["bqkv"][2 * Cq:], (1, 1, C), (0, 0, 1))
    if training:
        bdt = BF16 if C % 64 == 0 else F32
        kp = (nq + 63) // 64 * 64 if C % 64 == 0 else nq      # K of the dgrad GEMM, zero-padded to the UMMA K block
        wdq = torch.zeros((C, kp), dtype=bdt, device=dev)      # [c, (q | k | v | pad)] = [Wq ; Wk ; Wv]^T (persistent)
        plan.add(bp.Wq.detach(), wdq[:, :Cq], (C, 1, Cq), (1, 0, C), ld_dst=kp)
        plan.add(bp.Wk.detach(), wdq[:, Cq:2 * Cq], (C, 1, Cq), (1, 0, C), ld_dst=kp)
        plan.add(bp.Wv.detach(), wdq[:, 2 * Cq:nq], (C, 1, C), (1, 0, C), ld_dst=kp)
        pk["wdqkv"] = wdq


class NetPacks:
    """Persistent packed operands of a whole UNetDFCSA + the plan that refreshes them (cached on the module)."""

    def __init__(self, net, keep, fold=False):
        dev = next(net.parameters()).device
        self.sig = NetPacks.signature(net)
        self.plan = PackPlan(dev)
        self.bps = [make_block_params(b) for b in _blocks(net)]
        self.fold, self.affines, self.content = fold, [], None
        # fold: inference-mode operands with the running-statistics BatchNorms folded in (DFC-SA blocks; the ablation
        # block types keep the unfolded eval path)
        self.pks = [pack_block_weights_folded(bp, self.plan, self.affines) if (fold and bp.kind == "dfc") else
                    pack_block_weights(bp, keep, i > 0, self.plan) for i, bp in enumerate(self.bps)]
        self.up_fwd, self.up_bwd = [], []
        for up in (net.up4, net.up3, net.up2, net.up1):
            Ci_t, Co_t = up.weight.shape[0], up.weight.shape[1]
            tc = Ci_t % 64 == 0 and Co_t % 64 == 0
            wt = _e((4 * Co_t, Ci_t), F16 if tc else F32, dev)         # [(q, co), ci]
            self.plan.add(up.weight.detach(), wt, (4, Co_t, Ci_t), (1, 4, Co_t * 4))
            self.up_fwd.append(wt)
            if keep:
                wdt = _e((Ci_t, 4 * Co_t), BF16 if tc else F32, dev)   # [ci, (q, co)]
                self.plan.add(up.weight.detach(), wdt, (Ci_t, 4, Co_t), (Co_t * 4, 1, 4))
                self.up_bwd.append(wdt)
        fc = net.final_conv
        Cout, f0 = fc.weight.shape[0], fc.weight.shape[1]
        self.wdf = None
        if keep:
            self.wdf = _e((f0, Cout), F32, dev)                         # final conv dgrad operand [f0, Cout]
            self.plan.add(fc.weight.detach(), self.wdf, (f0, 1, Cout), (1, 0, f0))

    @staticmethod
    def signature(net):
        return tuple(p.data_ptr() for p in net.parameters())

    @staticmethod
    def content_signature(net):
        """changes whenever parameters or BatchNorm buffers may have changed: torch's in-place version counters (optimizer
        steps of torch.optim, load_state_dict, manual edits) + WEIGHTS_EPOCH for writes through raw pointers (FusedSGD, the
        running-statistics update of a training forward)."""
        return (WEIGHTS_EPOCH, sum(t._version for t in net.parameters()) + sum(t._version for t in net.buffers()))

    @staticmethod
    def get(net, keep, fold=False):
        cache = net.__dict__.setdefault("_dfcsa_packs", {})
        key = (keep, fold)
        pk = cache.get(key)
        if pk is None or pk.sig != NetPacks.signature(net):
            pk = cache[key] = NetPacks(net, keep, fold)
        if fold:        # inference: weights rarely change between calls - re-fold only when they did
            content = NetPacks.content_signature(net)
            if pk.content != content:
                for bn, conv_bias, scale, shift in pk.affines:
                    ops.bn_eval_affine(bn.weight.detach(), bn.bias.detach(), conv_bias.detach() if conv_bias is not None else None,
                                       bn.running_mean, bn.running_var, BN_EPS, scale, shift)
                pk.plan.run()
                pk.content = content
        else:
            pk.plan.run()
        return pk


class BlockCtx:
    pass


def _pool_side(bp, H, W):
    """Side of the pooled attention map.  pool_size=None is the full-resolution attention of ablation 3 (reference
    models/unet_dfc_sa_ablation_attention.py:15-26): attention over all H*W positions, which is exactly the pooled
    block with pool_size = H (adaptive_avg_pool2d to the input size and bilinear re-sizing to the same size are both
    the identity), so the same kernels run it."""
    if bp.P is not None:
        return bp.P
    if H != W:
        raise NotImplementedError("dfcsa: full-resolution attention is implemented for square feature maps (the 224/512/1024 "
                                  "hot-path shapes)")
    return H


def _bn_affine(bn, conv_bias, s_sum, s_sq, count, training, dev):
    Cn = bn.weight.numel()
    aff = _e((4, Cn), F32, dev)   # scale, shift, mean, invstd
    if training:
        ops.bn_finalize(s_sum, s_sq, count, bn.weight.detach(), bn.bias.detach(), conv_bias.detach() if conv_bias is not None else None,
                        bn.running_mean, bn.running_var, BN_MOMENTUM, BN_EPS, aff[0], aff[1], aff[2], aff[3])
    else:
        ops.bn_eval_affine(bn.weight.detach(), bn.bias.detach(), conv_bias.detach() if conv_bias is not None else None,
                           bn.running_mean, bn.running_var, BN_EPS, aff[0], aff[1])
    return aff


def _conv_bn(B, H, W, segs, w, N, out, st, bn, conv_bias, training, dev, **kw):
    """conv_gemm followed by the BatchNorm affine of its first bn.num_features output columns (scale, shift, mean, invstd).
    Training on the tcgen05 backend: the finalize is folded into the GEMM launch (its last CTA computes it, dfcsa_bn_fold_t);
    SIMT backend: a dfcsa_bn_finalize launch; eval: running statistics."""
    backend = _backend(segs, w, N, out)
    Cn = bn.weight.numel()
    M = B * H * W
    if training and backend == BACKEND_TC and _BN_FOLD:
        aff = _e((4, Cn), F32, dev)
        fold = ops.bn_fold(bn.weight.detach(), bn.bias.detach(), conv_bias.detach() if conv_bias is not None else None,
                           bn.running_mean, bn.running_var, BN_MOMENTUM, BN_EPS, M, aff, _z((1,), torch.int32, dev))
        ops.conv_gemm(B, H, W, segs, w, N, out, stats=st, backend=backend, bn=fold, **kw)
        return aff
    ops.conv_gemm(B, H, W, segs, w, N, out, stats=st if training else None, backend=backend, **kw)
    return _bn_affine(bn, conv_bias, st[0:Cn] if training else None, st[N:N + Cn] if training else None, M, training, dev)


_BN_FOLD = _os.environ.get("DFCSA_BN_FOLD", "1") != "0"     # A/B switch: "0" = separate dfcsa_bn_finalize launches
_ATTN_SAVE_BYTES = 4 << 30     # keep softmax(q k^T) for backward only below this size; above it backward recomputes it
_ATTN_CHUNK_BYTES = 12 << 30   # images are processed in chunks so that one [chunk, N, N] fp32 buffer stays below this


def _attn_chunk(B, N):
    return max(1, min(B, _ATTN_CHUNK_BYTES // (N * N * 4)))


def _attn_probs(qkv, nb, N, Cq, nq, out, lse=None, have_lse=False):
    """out[b] = softmax_j(q_i . k_j) for nb images; qkv: [nb*N, nq] rows (q | k | v).  fp32 qkv -> fp32 SIMT product;
    fp16 qkv -> tcgen05 batched GEMM (out is then usually fp16 too).  lse: optional [nb*N] fp32 that receives the row
    log-sum-exp (have_lse: it already holds it - the backward recompute - and the statistics pass is skipped)."""
    if qkv.dtype != F32:
        # two tensor-core passes over q k^T (K = Cq is tiny): row log-sum-exp, then exp(s - lse) stored directly as the
        # 16-bit probabilities - the [N, N] fp32 logits never exist in HBM
        ops.softmax_bgemm(nb, N, N, Cq, qkv[:, :Cq], N * nq, nq, qkv[:, Cq:2 * Cq], N * nq, nq, out, lse=lse, have_lse=have_lse)
        return
    S = _e((nb, N, N), F32, qkv.device)
    ops.sgemm(nb, N, N, Cq, qkv[:, :Cq], (N * nq, nq, 1), qkv[:, Cq:2 * Cq], (N * nq, 1, nq), S, (N * N, N, 1))
    ops.softmax_rows(S, out)


_WINDOW_TERMS = _os.environ.get("DFCSA_OLD_REDUCE2", "0") != "1"    # A/B switch for measurements (old: gather pass branch_bwd_reduce2)
_ATTN_FUSED = True       # forward attention without the N^2 round trip when the probabilities are not kept for backward
_ATTN_FUSED_BWD = True   # ... and the backward without any N^2 tensor (dfcsa_attn_bwd_fused)
_ATTN_SMALL_MAX_N = 32   # up to here the whole attention core is one fp32 kernel per direction (dfcsa_attn_small_*)
_ATTN_TC_MIN_N = 64      # below this the attention products are a few KFLOP per image: fp32 FMA, no tensor cores


def _attn_tc(bp, N):
    return bp.C % 64 == 0 and N >= _ATTN_TC_MIN_N and N % 8 == 0


def attention_forward(bp, pk, pooled, B, N, ctx=None):
    """q/k/v 1x1 convs (ONE tensor-core GEMM over [Wq ; Wk ; Wv]), softmax(q k^T), attn v on the pooled map
    [B*N, C] (reference :28-34).  pooled is fp32; the GEMM reads an fp16 copy.  For N >= 64 (pool_size >= 8 and
    the full-resolution attention of ablation 3) q k^T and attn v run as tcgen05 batched GEMMs on fp16 copies of
    q / k / v with the probabilities stored in fp16."""
    dev = pooled.device
    C, Cq = bp.C, bp.Cq
    BN = B * N
    nq = 2 * Cq + C
    tc = C % 64 == 0
    tca = _attn_tc(bp, N)
    p16 = pooled
    if tc:
        p16 = _e((BN, C), F16, dev)
        ops.cast2d(pooled, p16)
    qkv = _e((BN, nq), F32, dev)
    segs = [(p16, TAP_1x1)]
    ops.conv_gemm(1, 1, BN, segs, pk["wqkv"], nq, qkv, bias=pk["bqkv"], backend=_backend(segs, pk["wqkv"], nq, qkv))
    qsrc = qkv
    if tca:
        qsrc = _e((BN, nq), F16, dev)
        ops.cast2d(qkv, qsrc)
    if N <= _ATTN_SMALL_MAX_N:          # one kernel per image batch: S, softmax and PV on the 16x16 maps of pool_size 4
        attn = _e((B, N, N), F32, dev)
        o = _e((B, N, C), F32, dev)
        ops.attn_small_fwd(qkv, B, N, Cq, C, attn, o)
        if ctx is not None:
            ctx.qkv, ctx.attn, ctx.qkv16 = qkv, attn, None
            ctx.pooled_w = p16                             # weight-gradient operand (fp16, or fp32 on the SIMT path)
        return o
    # softmax(q k^T) v, a few images at a time when [N, N] is large (full-resolution attention: N = H*W)
    adt = F16 if tca else F32
    # exp + P V in one kernel whenever the shape allows: the probabilities are then never stored in the forward pass and the
    # backward rebuilds them (straight to bf16, the type it needs) from the saved row log-sum-exp
    fused = tca and C <= 128 and 8 <= Cq <= 64 and _ATTN_FUSED
    keep_attn = ctx is not None and not fused and B * N * N * (2 if tca else 4) <= _ATTN_SAVE_BYTES
    attn = _e((B, N, N), adt, dev) if keep_attn else None
    o = _e((B, N, C), F32, dev)
    # probabilities not kept: the backward recomputes them from the row log-sum-exp saved here (one GEMM pass, not two)
    lse = _e((BN,), F32, dev) if (tca and not keep_attn) else None
    if fused:
        # row statistics, then exp + P V in one kernel: the [N, N] probabilities never reach HBM in the forward pass
        ops.attn_row_lse(B, N, Cq, qsrc, nq, lse)
        ops.attn_pv_fused(qsrc, B, N, Cq, C, lse, o)
    ch = _attn_chunk(B, N)
    for b0 in range(0, 0 if fused else B, ch):
        nb = min(ch, B - b0)
        rows = slice(b0 * N, (b0 + nb) * N)
        A = attn[b0:b0 + nb] if keep_attn else _e((nb, N, N), adt, dev)
        _attn_probs(qsrc[rows], nb, N, Cq, nq, A, lse=lse[rows] if lse is not None else None)
        v = qsrc[rows, 2 * Cq:]
        if tca:     # o[i,c] = sum_j attn[i,j] v[j,c]: A K-major, V read MN-major (rows = keys)
            ops.bgemm(nb, N, C, N, A, N * N, N, False, v, N * nq, nq, True, o[b0:b0 + nb], N * C, C)
        else:
            ops.sgemm(nb, N, C, N, A, (N * N, N, 1), v, (N * nq, nq, 1), o[b0:b0 + nb], (N * C, C, 1))
    if ctx is not None:
        ctx.qkv, ctx.attn, ctx.qkv16, ctx.lse = qkv, attn, (qsrc if tca else None), lse
        ctx.pooled_w = p16                                 # weight-gradient operand (fp16, or fp32 on the SIMT path)
    return o


def attention_backward(bp, pk, ctx, d_o, B, N, grads):
    """Backward of attention_forward; returns dpooled [B*N, C] fp32 and fills the q/k/v weight and bias gradients."""
    dev = d_o.device
    C, Cq = bp.C, bp.Cq
    BN = B * N
    nq = 2 * Cq + C
    tca = _attn_tc(bp, N)
    qkv, attn = ctx.qkv, ctx.attn
    dqkv = _e((BN, nq), F32, dev)
    if tca:
        qkvb, dob = _e((BN, nq), BF16, dev), _e((BN, C), BF16, dev)     # gradient-side operands are bf16
        ops.cast2d(qkv, qkvb)
        ops.cast2d(d_o, dob)
        Drow = _e((BN,), F32, dev)       # D_i = sum_j dP_ij P_ij = sum_c dO_ic O_ic: makes the softmax backward one pass
        ops.rowdot(d_o, ctx.o.view(BN, C), Drow)
    ch = _attn_chunk(B, N)
    small = N <= _ATTN_SMALL_MAX_N
    if small:          # the kernel also accumulates the q / k / v bias gradients (column sums of dq / dk / dv)
        ops.attn_small_bwd(qkv, attn, d_o, B, N, Cq, C, dqkv, grads[bp.bq], grads[bp.bk], grads[bp.bv])
    fused = (not small) and tca and attn is None and ctx.lse is not None and C <= 128 and Cq <= 32 and _ATTN_FUSED_BWD
    if fused:       # dq | dk | dv without a single [N, N] tensor in HBM: P and dS are rebuilt on chip, tile by tile
        ops.attn_bwd_fused(ctx.qkv16, qkvb, dob, B, N, Cq, C, ctx.lse, Drow, dqkv)
    for b0 in range(0, 0 if (small or fused) else B, ch):
        nb = min(ch, B - b0)
        rows = slice(b0 * N, (b0 + nb) * N)
        dq, dk, dv = dqkv[rows, :Cq], dqkv[rows, Cq:2 * Cq], dqkv[rows, 2 * Cq:]
        if attn is not None:
            A = attn[b0:b0 + nb]
        elif tca:                              # not saved (too large): recompute the probabilities, straight to bf16
            A = _e((nb, N, N), BF16, dev)
            _attn_probs(ctx.qkv16[rows], nb, N, Cq, nq, A, lse=ctx.lse[rows], have_lse=True)
        else:
            A = _e((nb, N, N), F32, dev)
            _attn_probs(qkv[rows], nb, N, Cq, nq, A)
        if tca:
            q, k, v = qkvb[rows, :Cq], qkvb[rows, Cq:2 * Cq], qkvb[rows, 2 * Cq:]
            do = dob[rows]
            if A.dtype == BF16:
                Ab = A
            else:
                Ab = _e((nb, N, N), BF16, dev)
                ops.cast2d(A.view(nb * N, N), Ab.view(nb * N, N))
            ops.bgemm(nb, N, C, N, Ab, N * N, N, True, do, N * C, C, True, dv, N * nq, nq)         # dv[j,c] = sum_i attn[i,j] do[i,c]
            # dS = P * (dO V^T - D), the softmax backward in the epilogue of the dP product, written over P (bf16: a
            # gradient like every other activation gradient); dP itself never exists in HBM
            dS = Ab
            ops.softmax_bwd_bgemm(nb, N, N, C, do, N * C, C, v, N * nq, nq, Ab, Drow[rows], dS)
            del Ab
            ops.bgemm(nb, N, Cq, N, dS, N * N, N, False, k, N * nq, nq, True, dq, N * nq, nq)      # dq[i,c] = sum_j dS[i,j] k[j,c]
            ops.bgemm(nb, N, Cq, N, dS, N * N, N, True, q, N * nq, nq, True, dk, N * nq, nq)       # dk[j,c] = sum_i dS[i,j] q[i,c]
        else:
            q, k, v = qkv[rows, :Cq], qkv[rows, Cq:2 * Cq], qkv[rows, 2 * Cq:]
            do = d_o[rows]
            ops.sgemm(nb, N, C, N, A, (N * N, 1, N), do, (N * C, C, 1), dv, (N * nq, nq, 1))       # dv[j,c] = sum_i attn[i,j] do[i,c]
            dattn = _e((nb, N, N), F32, dev)
            ops.sgemm(nb, N, N, C, do, (N * C, C, 1), v, (N * nq, 1, nq), dattn, (N * N, N, 1))     # dattn[i,j] = sum_c do[i,c] v[j,c]
            dS = _e((nb, N, N), F32, dev)
            ops.softmax_rows_bwd(A, dattn, dS)
            del dattn
            ops.sgemm(nb, N, Cq, N, dS, (N * N, N, 1), k, (N * nq, nq, 1), dq, (N * nq, nq, 1))    # dq[i,c] = sum_j dS[i,j] k[j,c]
            ops.sgemm(nb, N, Cq, N, dS, (N * N, 1, N), q, (N * nq, nq, 1), dk, (N * nq, nq, 1))    # dk[j,c] = sum_i dS[i,j] q[i,c]
        del dS, A
    dq, dk, dv = dqkv[:, :Cq], dqkv[:, Cq:2 * Cq], dqkv[:, 2 * Cq:]
    if not small:
        for b, d in ((bp.bq, dq), (bp.bk, dk), (bp.bv, dv)):
            ops.colsum(d, grads[b])
    # dpooled = dqkv . [Wq ; Wk ; Wv]  and the three weight gradients, on the tensor cores when C allows
    wdq = pk["wdqkv"]
    kp = wdq.shape[1]
    dp = _e((BN, C), F32, dev)
    if wdq.dtype == BF16:
        d16 = _z((BN, kp), BF16, dev) if kp != nq else _e((BN, kp), BF16, dev)
        ops.cast2d(dqkv, d16[:, :nq])
    else:
        d16 = dqkv
    segs = [(d16, TAP_1x1)]
    ops.conv_gemm(1, 1, BN, segs, wdq, C, dp, backend=_backend(segs, wdq, C, dp))
    for W, lo, hi in ((bp.Wq, 0, Cq), (bp.Wk, Cq, 2 * Cq), (bp.Wv, 2 * Cq, nq)):
        _wgrad(1, 1, BN, ctx.pooled_w, TAP_1x1, d16[:, lo:hi], TAP_1x1, grads[W].view(hi - lo, C))
    return dp


def block_forward(bp, pk, x, B, H, W, y, yp=None, training=True, save=True):
    """One DynamicFusionConvAttnBlock.  x: [M, Ci] (fp16, or fp32 for the image); y / yp: fp16 output views (full
    resolution / 2x2 max-pooled).  Returns the context the backward needs."""
    if bp.kind == "local":
        return _local_block_forward(bp, pk, x, B, H, W, y, yp, training, save)
    if bp.kind != "dfc":
        return _branch_block_forward(bp, pk, x, B, H, W, y, yp, training, save)
    if pk.get("folded"):
        return _dfc_block_forward_folded(bp, pk, x, B, H, W, y, yp)
    dev = x.device
    C, Ci, P = bp.C, bp.Ci, _pool_side(bp, H, W)
    M = B * H * W
    ctx = BlockCtx() if (training and save) else None
    st = _z((10 * C,), F64, dev) if training else None
    L0, AR = _e((M, C), F16, dev), _e((M, 2 * C), F16, dev)
    segs3, segs1 = [(x, TAP_3x3)], [(x, TAP_1x1)]
    bn1 = _conv_bn(B, H, W, segs3, pk["w1"], C, L0, st[0:2 * C] if training else None, bp.bn1, bp.b1, training, dev)
    bn2 = _conv_bn(B, H, W, segs1, pk["w25"], 2 * C, AR, st[2 * C:6 * C] if training else None, bp.bn2, bp.b2, training, dev,
                   stats_cols=C)                                                       # only A0 feeds a BatchNorm
    A0, R = AR[:, :C], AR[:, C:]
    # pooled self-attention (when the backward will run, the pooling pass also emits the window means of the ReLU mask)
    # (only for coarse pooled maps, P <= 8: the two extra planes of the separable pooling pass cost [B, H, P, C] fp32 each,
    # which at P = 16 / 32 is more than the gather pass they replace - measured: P32 step 48.2 -> 50.4 ms with them)
    masks = ctx is not None and _WINDOW_TERMS and P <= 8
    tmp = _e((3 if masks else 1, B, H, P, C), F32, dev)
    pooled3 = _e((3 if masks else 1, B * P * P, C), F32, dev)
    pooled = pooled3[0]
    ops.bnrelu_pool_fwd(A0, B, H, W, bn2[0], bn2[1], P, tmp, pooled3, with_masks=masks)
    o = attention_forward(bp, pk, pooled, B, P * P, ctx)
    z = _e((M, 3 * C), F16, dev)
    ops.branch_act_fwd(L0, A0, B, H, W, bn1[0], bn1[1], bn2[0], bn2[1], o, P, bp.gamma.detach(), z, None)
    # gate
    G0 = _e((M, C), F16, dev)
    zLA = z[:, C:]
    segs = [(zLA, TAP_1x1)]
    bn3 = _conv_bn(B, H, W, segs, pk["w3"], C, G0, st[6 * C:8 * C] if training else None, bp.bn3, bp.b3, training, dev)
    ops.gate_mix_fwd(G0, bn3[0], bn3[1], z, None)
    # fusion
    F0 = _e((M, C), F16, dev)
    segs = [(z, TAP_1x1)]
    bn4 = _conv_bn(B, H, W, segs, pk["w4"], C, F0, st[8 * C:10 * C] if training else None, bp.bn4, bp.b4, training, dev)
    ops.block_out_fwd(F0, R, B, H, W, bn4[0], bn4[1], bp.res_scale.detach(), y, yp, None, None)
    if ctx is not None:
        ctx.B, ctx.H, ctx.W = B, H, W
        ctx.L0, ctx.A0, ctx.R, ctx.G0, ctx.F0, ctx.z, ctx.y, ctx.o = L0, A0, R, G0, F0, z, y, o
        ctx.bn1, ctx.bn2, ctx.bn3, ctx.bn4 = bn1, bn2, bn3, bn4
        ctx.win_means = pooled3[1:] if masks else None
    return ctx


_EVAL_EPI = _os.environ.get("DFCSA_EVAL_EPI", "1") != "0"     # A/B switch: "0" = separate gate-mix / residual passes


def _dfc_block_forward_folded(bp, pk, x, B, H, W, y, yp):
    """Inference-mode DynamicFusionConvAttnBlock with the BatchNorms folded into the GEMMs (pack_block_weights_folded):
    4 GEMMs with bias (+ ReLU) epilogues and 4 streaming passes (pool 1 E, A = gamma * up(o) + a 2 E, gate mix 4 E, residual
    sum + max pool 3.25 E); no statistics, no affine passes, no pre-BatchNorm tensors (L0 / A0 / F0 never exist)."""
    dev = x.device
    C, P = bp.C, _pool_side(bp, H, W)
    M = B * H * W
    ones, zeros = pk["ones"], pk["zeros"]
    z = _e((M, 3 * C), F16, dev)            # [f | L | A]
    AR = _e((M, 2 * C), F16, dev)           # [a | R]
    L = z[:, C:2 * C]
    segs3, segs1 = [(x, TAP_3x3)], [(x, TAP_1x1)]
    ops.conv_gemm(B, H, W, segs3, pk["w1"], C, L, bias=pk["t1"], act=1, backend=_backend(segs3, pk["w1"], C, L))
    ops.conv_gemm(B, H, W, segs1, pk["w25"], 2 * C, AR, bias=pk["b25"], act=1, act_cols=C, backend=_backend(segs1, pk["w25"], 2 * C, AR))
    a, R = AR[:, :C], AR[:, C:]
    tmp = _e((B, H, P, C), F32, dev)
    pooled = _e((B * P * P, C), F32, dev)
    ops.bnrelu_pool_fwd(a, B, H, W, ones, zeros, P, tmp, pooled)        # a >= 0 already: the fused affine + ReLU is the identity
    o = attention_forward(bp, pk, pooled, B, P * P, None)
    ops.branch_act_fwd(None, a, B, H, W, None, None, ones, zeros, o, P, bp.gamma.detach(), z, None)   # A = gamma * up(o) + a
    segs = [(z[:, C:], TAP_1x1)]
    f = z[:, :C]
    if _EVAL_EPI and C % 32 == 0 and _backend(segs, pk["w3"], C, f) == BACKEND_TC:
        # f = sigmoid(G) L + (1 - sigmoid(G)) A from the gate conv's epilogue (L, A: the rows its own TMA loads just pulled
        # through L2): the gate logits never reach HBM and the 4 E mixing pass is gone
        ops.conv_gemm(B, H, W, segs, pk["w3"], C, f, bias=pk["t3"], epi=("gate_mix", z[:, C:2 * C], z[:, 2 * C:]))
    else:
        G = _e((M, C), F16, dev)
        ops.conv_gemm(B, H, W, segs, pk["w3"], C, G, bias=pk["t3"], backend=_backend(segs, pk["w3"], C, G))
        ops.gate_mix_fwd(G, ones, zeros, z, None)                        # f = sigmoid(G) L + (1 - sigmoid(G)) A
    segs = [(z, TAP_1x1)]
    if _EVAL_EPI and yp is None and C % 32 == 0 and y.dtype == F16 and _backend(segs, pk["w4"], C, y) == BACKEND_TC:
        # blocks without a max pool behind them (decoder, bottleneck): y = relu(F) + res_scale * R from the fusion conv's
        # epilogue; F never reaches HBM
        ops.conv_gemm(B, H, W, segs, pk["w4"], C, y, bias=pk["t4"], act=1, epi=("residual", R, bp.res_scale.detach().reshape(1)))
    else:
        F = _e((M, C), F16, dev)
        ops.conv_gemm(B, H, W, segs, pk["w4"], C, F, bias=pk["t4"], act=1, backend=_backend(segs, pk["w4"], C, F))
        ops.sum_out_fwd(F, None, R, B, H, W, bp.res_scale.detach(), y, yp)   # y = F + res_scale * R (+ 2x2 max pool)
    return None


def _local_block_forward(bp, pk, x, B, H, W, y, yp, training, save):
    """LocalOnlyBlock: the DFC block's conv branch feeding the DFC block's output stage directly."""
    dev = x.device
    C, M = bp.C, B * H * W
    ctx = BlockCtx() if (training and save) else None
    st = _z((2 * C,), F64, dev) if training else None
    L0, R = _e((M, C), F16, dev), _e((M, C), F16, dev)
    segs3, segs1 = [(x, TAP_3x3)], [(x, TAP_1x1)]
    ops.conv_gemm(B, H, W, segs3, pk["w1"], C, L0, stats=st, backend=_backend(segs3, pk["w1"], C, L0))
    ops.conv_gemm(B, H, W, segs1, pk["w5"], C, R, backend=_backend(segs1, pk["w5"], C, R))
    bn1 = _bn_affine(bp.bn1, bp.b1, st[0:C] if training else None, st[C:2 * C] if training else None, M, training, dev)
    ops.block_out_fwd(L0, R, B, H, W, bn1[0], bn1[1], bp.res_scale.detach(), y, yp, None, None)
    if ctx is not None:
        ctx.B, ctx.H, ctx.W = B, H, W
        ctx.L0, ctx.R, ctx.y, ctx.bn1 = L0, R, y, bn1
    return ctx


def _local_block_backward(bp, pk, ctx, xw, dskip, dyp, dx_out, grads):
    dev = dskip.device if dskip is not None else dyp.device
    B, H, W = ctx.B, ctx.H, ctx.W
    C, Ci = bp.C, bp.Ci
    M = B * H * W
    red = _z((8 * C + 2,), F64, dev)       # the layout dfcsa_block_param_grads reads: [red1 | - | - | - | d res_scale | -]
    red1, drs = red[0:2 * C], red[8 * C:8 * C + 1]
    bn1 = ctx.bn1
    dy = dskip if dskip is not None else _e((M, C), BF16, dev)
    # out = relu(bn1(L0)) + res_scale * R
    ops.block_out_bwd_reduce(dskip, dyp, ctx.y, ctx.L0, ctx.R, B, H, W, bn1[0], bn1[1], bn1[2], bn1[3], dy, red1, drs)
    dL0 = _e((M, C), BF16, dev)
    ops.bn_bwd_apply(dy, ctx.L0, bn1[0], bn1[1], bn1[2], bn1[3], red1, 0, dL0)
    ops.block_param_grads(red, C, [(grads[bp.bn1.weight], grads[bp.bn1.bias]), None, None, None], drs=grads[bp.res_scale])
    if dx_out is not None:
        segs = [(dL0, TAP_3x3), (dy, TAP_1x1)]
        ops.conv_gemm(B, H, W, segs, pk["wd15"], Ci, dx_out, backend=_backend(segs, pk["wd15"], Ci, dx_out))
    dW1p = _z((C, 9 * Ci), F32, dev)
    _wgrad(B, H, W, xw, TAP_3x3, dL0, TAP_1x1, dW1p)
    ops.permute3(dW1p, grads[bp.W1], (C, Ci, 9), (9 * Ci, 1, Ci))
    if bp.W5 is not None:
        _wgrad(B, H, W, xw, TAP_1x1, dy, TAP_1x1, grads[bp.W5].view(C, Ci), alpha=bp.res_scale.detach().reshape(1))


def _branch_block_forward(bp, pk, x, B, H, W, y, yp, training, save):
    """AttentionOnly / AdditionFusion / ConcatFusion blocks: the DFC block's two branches without the gate.  The
    attention-only block has no conv branch; the two-branch kernels then run with the conv-branch operand aliased to the
    attention branch's (its outputs are ignored)."""
    dev = x.device
    C, P = bp.C, _pool_side(bp, H, W)
    M = B * H * W
    ctx = BlockCtx() if (training and save) else None
    st = _z((10 * C,), F64, dev) if training else None
    AR = _e((M, 2 * C), F16, dev)
    segs1 = [(x, TAP_1x1)]
    ops.conv_gemm(B, H, W, segs1, pk["w25"], 2 * C, AR, stats=st[2 * C:6 * C] if training else None,
                  backend=_backend(segs1, pk["w25"], 2 * C, AR))
    A0, R = AR[:, :C], AR[:, C:]
    bn2 = _bn_affine(bp.bn2, bp.b2, st[2 * C:3 * C] if training else None, st[4 * C:5 * C] if training else None, M, training, dev)
    if bp.has_l:
        L0 = _e((M, C), F16, dev)
        segs3 = [(x, TAP_3x3)]
        ops.conv_gemm(B, H, W, segs3, pk["w1"], C, L0, stats=st[0:2 * C] if training else None, backend=_backend(segs3, pk["w1"], C, L0))
        bn1 = _bn_affine(bp.bn1, bp.b1, st[0:C] if training else None, st[C:2 * C] if training else None, M, training, dev)
    else:
        L0, bn1 = A0, bn2
    tmp = _e((B, H, P, C), F32, dev)
    pooled = _e((B * P * P, C), F32, dev)
    ops.bnrelu_pool_fwd(A0, B, H, W, bn2[0], bn2[1], P, tmp, pooled)
    o = attention_forward(bp, pk, pooled, B, P * P, ctx)
    z = _e((M, 3 * C), F16, dev)            # [unused | L | A]
    ops.branch_act_fwd(L0, A0, B, H, W, bn1[0], bn1[1], bn2[0], bn2[1], o, P, bp.gamma.detach(), z, None)
    F0 = bn4 = None
    if bp.kind == "concat":
        F0 = _e((M, C), F16, dev)
        segs = [(z[:, C:], TAP_1x1)]
        ops.conv_gemm(B, H, W, segs, pk["w4c"], C, F0, stats=st[8 * C:10 * C] if training else None, backend=_backend(segs, pk["w4c"], C, F0))
        bn4 = _bn_affine(bp.bn4, bp.b4, st[8 * C:9 * C] if training else None, st[9 * C:10 * C] if training else None, M, training, dev)
        ops.block_out_fwd(F0, R, B, H, W, bn4[0], bn4[1], bp.res_scale.detach(), y, yp, None, None)
    else:
        ops.sum_out_fwd(z[:, 2 * C:], z[:, C:2 * C] if bp.has_l else None, R, B, H, W, bp.res_scale.detach(), y, yp)
    if ctx is not None:
        ctx.B, ctx.H, ctx.W = B, H, W
        ctx.L0, ctx.A0, ctx.R, ctx.F0, ctx.z, ctx.y, ctx.o = L0, A0, R, F0, z, y, o
        ctx.bn1, ctx.bn2, ctx.bn4 = bn1, bn2, bn4
    return ctx


def _branch_block_backward(bp, pk, ctx, xw, dskip, dyp, dx_out, grads):
    dev = dskip.device if dskip is not None else dyp.device
    B, H, W = ctx.B, ctx.H, ctx.W
    C, Ci, P = bp.C, bp.Ci, _pool_side(bp, H, W)
    M = B * H * W
    red = _z((8 * C + 2,), F64, dev)
    red1, red2, red4 = red[0:2 * C], red[2 * C:4 * C], red[6 * C:8 * C]
    drs, dgam = red[8 * C:8 * C + 1], red[8 * C + 1:8 * C + 2]
    bn1, bn2 = ctx.bn1, ctx.bn2
    dy = dskip if dskip is not None else _e((M, C), BF16, dev)
    dz = _e((M, 3 * C), BF16, dev)
    dLA = dz[:, C:]
    if bp.kind == "concat":
        bn4 = ctx.bn4
        ops.block_out_bwd_reduce(dskip, dyp, ctx.y, ctx.F0, ctx.R, B, H, W, bn4[0], bn4[1], bn4[2], bn4[3], dy, red4, drs)
        dF0 = _e((M, C), BF16, dev)
        ops.bn_bwd_apply(dy, ctx.F0, bn4[0], bn4[1], bn4[2], bn4[3], red4, 0, dF0)
        segs = [(dF0, TAP_1x1)]
        ops.conv_gemm(B, H, W, segs, pk["wd4c"], 2 * C, dLA, backend=_backend(segs, pk["wd4c"], 2 * C, dLA))
        _wgrad(B, H, W, ctx.z[:, C:], TAP_1x1, dF0, TAP_1x1, grads[bp.W4].view(C, 2 * C))
    else:
        # y = [L +] A + rs*R: dL = dA = dy.  The output-stage reduction also gathers the pooled gradient and d res_scale;
        # its BatchNorm sums (taken over y with bn2's constants here) are not used
        ops.block_out_bwd_reduce(dskip, dyp, ctx.y, ctx.y, ctx.R, B, H, W, bn2[0], bn2[1], bn2[2], bn2[3], dy, red4, drs)
        ops.cast2d(dy, dz[:, C:2 * C])
        ops.cast2d(dy, dz[:, 2 * C:])
    tmp = _e((B, H, P, C), F32, dev)
    d_o = _e((B * P * P, C), F32, dev)
    ops.branch_bwd_reduce1(dz, ctx.L0, None, B, H, W, bn1[0], bn1[1], bn1[2], bn1[3], None, None, ctx.o, P,
                           bp.gamma.detach(), red1, dgam, tmp, d_o)
    dpooled = attention_backward(bp, pk, ctx, d_o, B, P * P, grads)
    ops.branch_bwd_reduce2(dz, ctx.A0, B, H, W, bn2[0], bn2[1], bn2[2], bn2[3], dpooled, P, red2)
    dL0, dA0 = _e((M, C), BF16, dev), _e((M, C), BF16, dev)
    ops.branch_bwd_apply(dz, ctx.L0, ctx.A0, B, H, W, bn1, red1, bn2, red2, dpooled, P, dL0, dA0)
    ops.block_param_grads(red, C, [(grads[bp.bn1.weight], grads[bp.bn1.bias]) if bp.has_l else None,
                                   (grads[bp.bn2.weight], grads[bp.bn2.bias]), None,
                                   (grads[bp.bn4.weight], grads[bp.bn4.bias]) if bp.kind == "concat" else None],
                          drs=grads[bp.res_scale], dgam=grads[bp.gamma])
    if dx_out is not None:
        segs = ([(dL0, TAP_3x3)] if bp.has_l else []) + [(dA0, TAP_1x1), (dy, TAP_1x1)]
        ops.conv_gemm(B, H, W, segs, pk["wd125"], Ci, dx_out, backend=_backend(segs, pk["wd125"], Ci, dx_out))
    if bp.has_l:
        dW1p = _z((C, 9 * Ci), F32, dev)
        _wgrad(B, H, W, xw, TAP_3x3, dL0, TAP_1x1, dW1p)
        ops.permute3(dW1p, grads[bp.W1], (C, Ci, 9), (9 * Ci, 1, Ci))
    if bp.W5 is not None:
        _wgrad(B, H, W, xw, TAP_1x1, dy, TAP_1x1, grads[bp.W5].view(C, Ci), alpha=bp.res_scale.detach().reshape(1),
               second=(dA0, grads[bp.W2].view(C, Ci), 0, None))
    else:
        _wgrad(B, H, W, xw, TAP_1x1, dA0, TAP_1x1, grads[bp.W2].view(C, Ci))


def _wgrad(B, H, W, x, x_mode, dy, dy_mode, dw2d, alpha=None, second=None):
    """second = (dy2, dw2, c_begin2, alpha2): another 1x1 weight gradient over the same x, same launch."""
    tc = ops.wgrad_tc_eligible(x, dy) and (x.dtype == dy.dtype or (x.dtype == F16 and dy.dtype == BF16))
    if second is not None:
        dy2 = second[0]
        tc = tc and ops.wgrad_tc_eligible(x, dy2) and dy2.dtype == dy.dtype and dy.shape[1] % 64 == 0 and second[2] % 32 == 0
    ops.conv_wgrad(B, H, W, x, x_mode, dy, dy_mode, dw2d, alpha=alpha, backend=BACKEND_TC if tc else BACKEND_SIMT, second=second)


def block_backward(bp, pk, ctx, xw, dskip, dyp, dx_out, grads):
    """Backward of block_forward.  xw: the block input (fp16, or the fp32 image) as the weight-gradient operand.
    dskip: bf16 gradient w.r.t. y ([M, C] view, updated in place when dyp is given); dyp: bf16 gradient w.r.t. the
    max-pooled output or None; dx_out: bf16 [M, Ci] view that receives the input gradient, or None.
    grads: dict parameter -> fp32 gradient tensor (zero-initialised where the kernels accumulate)."""
    if bp.kind == "local":
        return _local_block_backward(bp, pk, ctx, xw, dskip, dyp, dx_out, grads)
    if bp.kind != "dfc":
        return _branch_block_backward(bp, pk, ctx, xw, dskip, dyp, dx_out, grads)
    dev = dskip.device if dskip is not None else dyp.device
    B, H, W = ctx.B, ctx.H, ctx.W
    C, Ci, P = bp.C, bp.Ci, _pool_side(bp, H, W)
    M = B * H * W
    red = _z((8 * C + 2,), F64, dev)
    red1, red2, red3, red4 = red[0:2 * C], red[2 * C:4 * C], red[4 * C:6 * C], red[6 * C:8 * C]
    drs, dgam = red[8 * C:8 * C + 1], red[8 * C + 1:8 * C + 2]
    bn1, bn2, bn3, bn4 = ctx.bn1, ctx.bn2, ctx.bn3, ctx.bn4
    if dskip is None:
        dskip_buf = _e((M, C), BF16, dev)
        dy = dskip_buf
    else:
        dy = dskip
    # out = relu(bn4(F0)) + res_scale * R
    ops.block_out_bwd_reduce(dskip, dyp, ctx.y, ctx.F0, ctx.R, B, H, W, bn4[0], bn4[1], bn4[2], bn4[3], dy, red4, drs)
    dF0 = _e((M, C), BF16, dev)
    ops.bn_bwd_apply(dy, ctx.F0, bn4[0], bn4[1], bn4[2], bn4[3], red4, 0, dF0)
    # fusion conv, part 1: df = dF0 . W4^T[:, 0:C]  (the gate-mix backward only needs df)
    dz = _e((M, 3 * C), BF16, dev)
    segs = [(dF0, TAP_1x1)]
    df = dz[:, :C]
    ops.conv_gemm(B, H, W, segs, pk["wd4f"], C, df, backend=_backend(segs, pk["wd4f"], C, df))
    # gate / mix
    ops.gate_mix_bwd_reduce(dz, ctx.z, ctx.G0, bn3[0], bn3[1], bn3[2], bn3[3], red3)
    dG0 = _e((M, C), BF16, dev)
    ops.gate_mix_bwd_apply(dz, ctx.z, ctx.G0, bn3[0], bn3[1], bn3[2], bn3[3], red3, dG0)
    # fusion conv part 2 + gate conv in ONE dgrad GEMM: [dL' | dA'] = [dF0 | dG0] . [W4^T[:, C:3C] ; W3^T]
    dLA = dz[:, C:]
    segs = [(dF0, TAP_1x1), (dG0, TAP_1x1)]
    ops.conv_gemm(B, H, W, segs, pk["wd43"], 2 * C, dLA, backend=_backend(segs, pk["wd43"], 2 * C, dLA))
    # fusion-conv and gate-conv weight gradients in one launch: both read z = [f | L | A] (the gate conv only L | A)
    _wgrad(B, H, W, ctx.z, TAP_1x1, dF0, TAP_1x1, grads[bp.W4].view(C, 3 * C),
           second=(dG0, grads[bp.W3].view(C, 2 * C), C, None))
    # branches (reduce1 first adds the gate-mix terms df*g / df*(1-g) into dL / dA in place)
    tmp = _e((B, H, P, C), F32, dev)
    d_o = _e((B * P * P, C), F32, dev)
    ops.branch_bwd_reduce1(dz, ctx.L0, ctx.G0, B, H, W, bn1[0], bn1[1], bn1[2], bn1[3], bn3[0], bn3[1], ctx.o, P,
                           bp.gamma.detach(), red1, dgam, tmp, d_o)
    dpooled = attention_backward(bp, pk, ctx, d_o, B, P * P, grads)
    if ctx.win_means is not None:
        # BatchNorm-2 sums without a gather pass: the dA part from a plain (dA, A0) reduction, the pool^T(dpooled) part
        # from the window means the forward pooling emitted (adaptive_avg_pool^T is linear)
        ops.bn_bwd_reduce(dz[:, 2 * C:], ctx.A0, bn2[0], bn2[1], bn2[2], bn2[3], red2)
        ops.pool_window_terms(dpooled, ctx.win_means, B, P, C, bn2[2], bn2[3], red2)
    else:
        ops.branch_bwd_reduce2(dz, ctx.A0, B, H, W, bn2[0], bn2[1], bn2[2], bn2[3], dpooled, P, red2)
    dL0, dA0 = dF0, dG0     # both dead after the GEMMs above: reuse their storage
    ops.branch_bwd_apply(dz, ctx.L0, ctx.A0, B, H, W, bn1, red1, bn2, red2, dpooled, P, dL0, dA0)
    # every small parameter gradient of the block (4 x BatchNorm weight / bias, gamma, res_scale) in one launch
    ops.block_param_grads(red, C, [(grads[bn.weight], grads[bn.bias]) for bn in (bp.bn1, bp.bn2, bp.bn3, bp.bn4)],
                          drs=grads[bp.res_scale], dgam=grads[bp.gamma])
    # the three convs that read x
    if dx_out is not None:
        segs = [(dL0, TAP_3x3), (dA0, TAP_1x1), (dy, TAP_1x1)]
        ops.conv_gemm(B, H, W, segs, pk["wd125"], Ci, dx_out, backend=_backend(segs, pk["wd125"], Ci, dx_out))
    dW1p = _z((C, 9 * Ci), F32, dev)
    _wgrad(B, H, W, xw, TAP_3x3, dL0, TAP_1x1, dW1p)
    ops.permute3(dW1p, grads[bp.W1], (C, Ci, 9), (9 * Ci, 1, Ci))
    # attention-branch 1x1 conv and residual 1x1 conv weight gradients in one launch over the block input
    if bp.W5 is not None:
        _wgrad(B, H, W, xw, TAP_1x1, dy, TAP_1x1, grads[bp.W5].view(C, Ci), alpha=bp.res_scale.detach().reshape(1),
               second=(dA0, grads[bp.W2].view(C, Ci), 0, None))
    else:
        _wgrad(B, H, W, xw, TAP_1x1, dA0, TAP_1x1, grads[bp.W2].view(C, Ci))
    # conv biases in front of a train-mode BatchNorm have exactly zero gradient: grads[...] stay zero


class NetCtx:
    pass


def _blocks(net):
    return [net.down1, net.down2, net.down3, net.down4, net.bottleneck, net.up_conv4, net.up_conv3, net.up_conv2, net.up_conv1]


def net_forward(net, x_nchw, training, save=True):
    """UNetDFCSA.forward (reference :161-204).  Returns (logits NCHW fp32, ctx).  training selects batch-statistics
    BatchNorm; save keeps what the backward needs (and writes the bf16 shadows)."""
    keep = training and save
    dev = x_nchw.device
    if not x_nchw.is_cuda:
        raise RuntimeError("dfcsa: the model runs on CUDA tensors only (no CPU fallback)")
    B, Cin, H, W = x_nchw.shape
    if H < 16 or W < 16:
        raise ValueError("dfcsa: the network pools four times; H and W must be at least 16")
    x_nchw = x_nchw.contiguous().float()
    packs = NetPacks.get(net, keep, fold=not training)
    global _ARENA, WEIGHTS_EPOCH
    if training:
        WEIGHTS_EPOCH += 1          # bn_finalize rewrites the running statistics through raw pointers
        _ARENA = _net_arena(net, "fwd", dev)
        _ARENA.begin()
    try:
        return _net_forward(net, x_nchw, training, keep, packs)
    finally:
        _ARENA = None


def _net_forward(net, x_nchw, training, keep, packs):
    dev = x_nchw.device
    B, Cin, H, W = x_nchw.shape
    bps, pks = packs.bps, packs.pks
    f = [bps[i].C for i in range(4)]
    ctx = NetCtx() if keep else None
    Hs = [H >> i for i in range(5)]      # MaxPool2d(2, 2) floors odd sides (reference :164-173)
    Ws = [W >> i for i in range(5)]
    Ms = [B * Hs[i] * Ws[i] for i in range(5)]
    x0 = _e((Ms[0], Cin), F32, dev)
    ops.nchw_to_nhwc(x_nchw, x0, B, Cin, H, W)
    cat = [_e((Ms[i], 2 * f[i]), F16, dev) for i in range(4)]      # [up | skip] per level
    bctx = [None] * 9
    xin = x0
    xs = []              # the input of every block: the weight-gradient operand of its three input convs
    for i in range(4):   # encoder
        yp = _e((Ms[i + 1], f[i]), F16, dev)
        xs.append(xin)
        bctx[i] = block_forward(bps[i], pks[i], xin, B, Hs[i], Ws[i], cat[i][:, f[i]:], yp, training, keep)
        xin = yp
    u = _e((Ms[4], 2 * f[3]), F16, dev)
    xs.append(xin)
    bctx[4] = block_forward(bps[4], pks[4], xin, B, Hs[4], Ws[4], u, None, training, keep)
    ups = [net.up4, net.up3, net.up2, net.up1]
    upk, uin = [], []
    for j, lvl in enumerate((3, 2, 1, 0)):   # decoder
        up = ups[j]
        Ci_t, Co_t = up.weight.shape[0], up.weight.shape[1]
        tc = Ci_t % 64 == 0 and Co_t % 64 == 0
        wt = packs.up_fwd[j]
        segs = [(u, TAP_1x1)]
        # the ConvT output is (2h, 2w); when a side of this level is odd (inputs that are not multiples of 16) the
        # reference re-sizes it to the skip tensor's size with bilinear interpolation (:180-181) - off the hot path
        resized = (2 * Hs[lvl + 1], 2 * Ws[lvl + 1]) != (Hs[lvl], Ws[lvl])
        dst = _e((4 * Ms[lvl + 1], f[lvl]), F16, dev) if resized else cat[lvl][:, :f[lvl]]
        ops.conv_gemm(B, Hs[lvl + 1], Ws[lvl + 1], segs, wt, 4 * Co_t, dst, out_mode=OUT_CONVT2x2, bias=up.bias.detach(),
                      backend=BACKEND_TC if (tc and ops.tc_eligible(segs, 4 * Co_t, dst)) else BACKEND_SIMT)
        if resized:
            ops.resize_bilinear(dst, B, 2 * Hs[lvl + 1], 2 * Ws[lvl + 1], cat[lvl][:, :f[lvl]], Hs[lvl], Ws[lvl])
        if keep:
            upk.append(packs.up_bwd[j])
            uin.append(u)
        un = _e((Ms[lvl], f[lvl]), F16, dev)
        xs.append(cat[lvl])
        bctx[5 + j] = block_forward(bps[5 + j], pks[5 + j], cat[lvl], B, Hs[lvl], Ws[lvl], un, None, training, keep)
        u = un
    # final 1x1 conv (Co = out_channels is tiny: fp32 SIMT path, logits in fp32)
    fc = net.final_conv
    Cout = fc.weight.shape[0]
    logits_nhwc = _e((Ms[0], Cout), F32, dev)
    ops.conv_gemm(B, H, W, [(u, TAP_1x1)], fc.weight.detach().view(Cout, -1), Cout, logits_nhwc, bias=fc.bias.detach(), backend=BACKEND_SIMT)
    if Cout == 1:
        logits = logits_nhwc.view(B, 1, H, W)
    else:
        logits = _e((B, Cout, H, W), F32, dev)
        ops.nhwc_to_nchw(logits_nhwc, logits, B, Cout, H, W)
    if training:
        torch._foreach_add_([m.num_batches_tracked for bp in bps for m in bp.bns()], 1)
    if keep:
        ctx.bps, ctx.pks, ctx.bctx, ctx.xs = bps, pks, bctx, xs
        ctx.upk, ctx.uin, ctx.u_last = upk, uin, u
        ctx.dims = (B, Cin, H, W, f, Hs, Ws, Ms)
        ctx.wdf = packs.wdf
    return logits, ctx


def eval_forward_graphed(net, x_nchw):
    """Inference forward replayed from a CUDA graph (one per input shape; opt-in with `net.eval_cuda_graph = True`): the
    ~150 launches of a folded eval forward cost more host time than GPU time at batch 1 (C5's p50 latency).  The first two
    calls per shape run eagerly, the third captures.  Folded operands live in persistent buffers, so when the weights
    change they are re-folded in place before the replay and the graph stays valid.  Returns a fresh tensor."""
    cache = net.__dict__.setdefault("_dfcsa_eval_graphs", {})
    key = (tuple(x_nchw.shape), str(x_nchw.device))
    e = cache.setdefault(key, [0, None, None, None])
    e[0] += 1
    if e[0] <= 2:
        return net_forward(net, x_nchw, False, save=False)[0]
    NetPacks.get(net, False, fold=True)          # re-fold (eagerly, outside the graph) if parameters / buffers changed
    if e[1] is None:
        e[2] = x_nchw.detach().float().contiguous().clone()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            e[3] = net_forward(net, e[2], False, save=False)[0]
        e[1] = g
    e[2].copy_(x_nchw, non_blocking=True)
    e[1].replay()
    return e[3].clone()


def net_param_list(net):
    return [p for p in net.parameters()]


def net_backward(net, ctx, dlogits_nchw, grads, after_stage=None):
    """Backward of net_forward: fills grads[param] (fp32, must be zero-initialised) for every parameter.
    after_stage(k), if given, is called when the k-th gradient bucket of trainer.reduce_buckets() is complete."""
    global _ARENA
    _ARENA = _net_arena(net, "bwd", dlogits_nchw.device)
    _ARENA.begin()
    try:
        return _net_backward(net, ctx, dlogits_nchw, grads, after_stage)
    finally:
        _ARENA = None


def _net_backward(net, ctx, dlogits_nchw, grads, after_stage):
    stage = [0]

    def done():
        if after_stage is not None:
            after_stage(stage[0])
        stage[0] += 1
    dev = dlogits_nchw.device
    B, Cin, H, W, f, Hs, Ws, Ms = ctx.dims
    bps, pks, bctx = ctx.bps, ctx.pks, ctx.bctx
    fc = net.final_conv
    Cout = fc.weight.shape[0]
    dl = _e((Ms[0], Cout), BF16, dev)
    ops.nchw_to_nhwc(dlogits_nchw.contiguous().float(), dl, B, Cout, H, W)
    # final conv backward (K = Cout is tiny: SIMT)
    du = _e((Ms[0], f[0]), BF16, dev)
    ops.conv_gemm(B, H, W, [(dl, TAP_1x1)], ctx.wdf, f[0], du, backend=BACKEND_SIMT)
    ops.conv_wgrad(B, H, W, ctx.u_last, TAP_1x1, dl, TAP_1x1, grads[fc.weight].view(Cout, f[0]), backend=BACKEND_SIMT)
    ops.colsum(dl, grads[fc.bias])
    ups = [net.up4, net.up3, net.up2, net.up1]
    dcat = [None] * 4
    for j, lvl in reversed(list(enumerate((3, 2, 1, 0)))):   # up_conv1 first
        dcat[lvl] = _e((Ms[lvl], 2 * f[lvl]), BF16, dev)
        block_backward(bps[5 + j], pks[5 + j], bctx[5 + j], ctx.xs[5 + j], du, None, dcat[lvl], grads)
        if lvl == 3:
            done()          # {up_conv4}
        # ConvTranspose2d backward
        up = ups[j]
        Ci_t, Co_t = up.weight.shape[0], up.weight.shape[1]
        dup = dcat[lvl][:, :f[lvl]]
        if (2 * Hs[lvl + 1], 2 * Ws[lvl + 1]) != (Hs[lvl], Ws[lvl]):       # backward of the bilinear re-size (:180-181)
            dup_t = _e((4 * Ms[lvl + 1], f[lvl]), BF16, dev)
            ops.resize_bilinear_bwd(dup, B, 2 * Hs[lvl + 1], 2 * Ws[lvl + 1], dup_t, Hs[lvl], Ws[lvl])
            dup = dup_t
        du = _e((Ms[lvl + 1], Ci_t), BF16, dev)
        segs = [(dup, TAP_2x2S2)]
        ops.conv_gemm(B, Hs[lvl + 1], Ws[lvl + 1], segs, ctx.upk[j], Ci_t, du, backend=_backend(segs, ctx.upk[j], Ci_t, du))
        dWp = _z((Co_t, 4 * Ci_t), F32, dev)
        _wgrad(B, Hs[lvl + 1], Ws[lvl + 1], ctx.uin[j], TAP_1x1, dup, TAP_2x2S2, dWp)
        ops.permute3(dWp, grads[up.weight], (Ci_t, Co_t, 4), (1, 4 * Ci_t, Ci_t))
        ops.colsum(dup, grads[up.bias])
        if lvl != 3:
            done()          # buckets: {final, up_conv1, up1}, {up_conv2, up2}, {up_conv3, up3}
    # bottleneck
    dyp = _e((Ms[4], f[3]), BF16, dev)
    done()                  # {up4}
    block_backward(bps[4], pks[4], bctx[4], ctx.xs[4], du, None, dyp, grads)
    done()                  # {bottleneck}
    for i in (3, 2, 1, 0):   # encoder
        dx = _e((Ms[i], bps[i].Ci), BF16, dev) if i > 0 else None
        block_backward(bps[i], pks[i], bctx[i], ctx.xs[i], dcat[i][:, f[i]:], dyp, dx, grads)
        dyp = dx
        if i != 1:
            done()          # {down4}, {down3}, {down2, down1}
