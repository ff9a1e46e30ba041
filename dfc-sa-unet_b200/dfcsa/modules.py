"""nn.Module mirror of the reference model classes (models/unet_dfc_sa_res.py:5-220).

Same class names, constructor arguments, sub-module names and therefore the same 343-entry state_dict (SURVEY.md
App. D) and the same random initialisation for a given torch seed: the standard torch.nn layers are used purely as
parameter containers (their forward is never called).  forward() goes through libdfcsa via dfcsa.engine; gradients
come back through one torch.autograd.Function per network (or per stand-alone block), so torch.optim.SGD and
clip_grad_norm_ of the reference Trainer keep working on ordinary fp32 nn.Parameters.
"""
import torch
import torch.nn as nn

from . import engine, ops


class _AttnParams:
    """engine-side view of a stand-alone attention module (the fields engine.attention_forward / _backward read)."""

    def __init__(self, att):
        self.att, self.gamma = att, att.gamma
        self.Wq, self.bq = att.query_conv.weight, att.query_conv.bias
        self.Wk, self.bk = att.key_conv.weight, att.key_conv.bias
        self.Wv, self.bv = att.value_conv.weight, att.value_conv.bias
        self.C, self.Cq, self.P = self.Wv.shape[0], self.Wq.shape[0], att.pool_size


class _AttnFunction(torch.autograd.Function):
    """LightSelfAttention.forward on its own (reference models/unet_dfc_sa_res.py:20-39): adaptive pool -> q/k/v 1x1 convs,
    softmax(q k^T), attn v (the same engine.attention_forward the fused block uses) -> gamma * bilinear_up(out) + x.
    Inside DynamicFusionConvAttnBlock the pool and the up-sample are fused into the block's streaming kernels; this
    stand-alone form uses the general dfcsa_adaptive_pool / dfcsa_resize_bilinear kernels (fp32 at the module boundary)."""

    @staticmethod
    def forward(ctx, x, att, need_grad, *params):
        if not x.is_cuda:
            raise RuntimeError("dfcsa: the attention module runs on CUDA tensors only (no CPU fallback)")
        B, C, H, W = x.shape
        dev = x.device
        ap = _AttnParams(att)
        if C != ap.C:
            raise ValueError(f"dfcsa: attention built for {ap.C} channels got {C}")
        P = ap.P if ap.P is not None else H
        if ap.P is None and H != W:
            raise NotImplementedError("dfcsa: full-resolution attention is implemented for square feature maps")
        M, N = B * H * W, P * P
        x2 = engine._e((M, C), torch.float32, dev)
        ops.nchw_to_nhwc(x.contiguous().float(), x2, B, C, H, W)
        pooled = engine._e((B * N, C), torch.float32, dev)
        ops.adaptive_pool(x2, B, H, W, P, pooled)                                   # :24
        plan, pk = engine.PackPlan(dev), {}
        engine._pack_attention(ap, pk, need_grad, plan)
        plan.run()
        actx = engine.BlockCtx() if need_grad else None
        o = engine.attention_forward(ap, pk, pooled, B, N, actx)                     # :28-34
        out2 = engine._e((M, C), torch.float32, dev)
        ops.resize_bilinear(o.view(B * N, C), B, P, P, out2, H, W, alpha=att.gamma.detach(), add=x2)     # :36-38
        out = engine._e((B, C, H, W), torch.float32, dev)
        ops.nhwc_to_nchw(out2, out, B, C, H, W)
        if need_grad:
            actx.o = o
        ctx.ap, ctx.pk, ctx.actx, ctx.dims, ctx.params = ap, pk, actx, (B, C, H, W, P), params
        return out

    @staticmethod
    def backward(ctx, dout):
        if ctx.actx is None:
            raise RuntimeError("dfcsa: backward through an attention module that ran in no-grad mode")
        B, C, H, W, P = ctx.dims
        ap, actx = ctx.ap, ctx.actx
        dev = dout.device
        M, N = B * H * W, P * P
        dy2 = engine._e((M, C), torch.float32, dev)
        ops.nchw_to_nhwc(dout.contiguous().float(), dy2, B, C, H, W)
        grads = {p: torch.zeros_like(p) for p in ctx.params}
        o2 = actx.o.view(B * N, C)
        # d gamma = sum dout * up(o)
        U = engine._e((M, C), torch.float32, dev)
        ops.resize_bilinear(o2, B, P, P, U, H, W)
        rd = engine._e((M, 1), torch.float32, dev)
        ops.rowdot(dy2, U, rd)
        ops.colsum(rd, grads[ap.gamma])
        d_o = engine._e((B * N, C), torch.float32, dev)
        ops.resize_bilinear_bwd(dy2, B, P, P, d_o, H, W, alpha=ap.gamma.detach())
        dpooled = engine.attention_backward(ap, ctx.pk, actx, d_o, B, N, grads)
        dx2 = engine._e((M, C), torch.float32, dev)
        ops.adaptive_pool_bwd(dpooled, B, H, W, P, dx2, add=dy2)
        dx = engine._e((B, C, H, W), torch.float32, dev)
        ops.nhwc_to_nchw(dx2, dx, B, C, H, W)
        return (dx, None, None) + tuple(grads[p] for p in ctx.params)


class LightSelfAttention(nn.Module):
    """reference models/unet_dfc_sa_res.py:5-39.  Inside a DFC-SA block the math runs fused in the block kernels; called on
    its own (the reference class is an ordinary module) it runs through _AttnFunction."""

    def __init__(self, channels, pool_size=8, ablation_on_qk_channels=8):
        super().__init__()
        self.pool_size = pool_size
        self.query_conv = nn.Conv2d(channels, channels // ablation_on_qk_channels, kernel_size=1)
        self.key_conv = nn.Conv2d(channels, channels // ablation_on_qk_channels, kernel_size=1)
        self.value_conv = nn.Conv2d(channels, channels, kernel_size=1)
        self.gamma = nn.Parameter(torch.zeros(1))

    def forward(self, x):
        params = list(self.parameters())
        need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
        return _AttnFunction.apply(x, self, need_grad, *params)


class FullResolutionAttention(nn.Module):
    """reference models/unet_dfc_sa_ablation_attention.py:7-26 (ablation 3): same parameters as LightSelfAttention with
    channels // 8 query/key channels and no pooling; pool_size=None tells the engine to attend over all H*W positions."""

    def __init__(self, channels, **kwargs):
        super().__init__()
        self.pool_size = None
        self.query_conv = nn.Conv2d(channels, channels // 8, kernel_size=1)
        self.key_conv = nn.Conv2d(channels, channels // 8, kernel_size=1)
        self.value_conv = nn.Conv2d(channels, channels, kernel_size=1)
        self.gamma = nn.Parameter(torch.zeros(1))

    def forward(self, x):
        params = list(self.parameters())
        need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
        return _AttnFunction.apply(x, self, need_grad, *params)


class _BlockFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, block, need_grad, *params):
        B, Ci, H, W = x.shape
        dev = x.device
        if not x.is_cuda:
            raise RuntimeError("dfcsa: the block runs on CUDA tensors only (no CPU fallback)")
        training = block.training
        keep = training and need_grad
        bp = engine.make_block_params(block)
        plan = engine.PackPlan(dev)
        pk = engine.pack_block_weights(bp, keep, True, plan)
        plan.run()
        M = B * H * W
        xin = engine._e((M, Ci), torch.float16 if bp.tc else torch.float32, dev)
        ops.nchw_to_nhwc(x.contiguous().float(), xin, B, Ci, H, W)
        xb = xin if keep else None      # the weight-gradient kernels read the fp16 (or fp32) input directly
        y = engine._e((M, bp.C), torch.float16, dev)
        bctx = engine.block_forward(bp, pk, xin, B, H, W, y, training=training, save=keep)
        out = engine._e((B, bp.C, H, W), torch.float32, dev)
        ops.nhwc_to_nchw(y, out, B, bp.C, H, W)
        ctx.block, ctx.bp, ctx.pk, ctx.bctx, ctx.xb, ctx.dims = block, bp, pk, bctx, xb, (B, Ci, H, W)
        ctx.params = params
        return out

    @staticmethod
    def backward(ctx, dout):
        B, Ci, H, W = ctx.dims
        bp = ctx.bp
        dev = dout.device
        if ctx.bctx is None:
            raise RuntimeError("dfcsa: backward through a block that ran in eval / no-grad mode")
        M = B * H * W
        dy = engine._e((M, bp.C), torch.bfloat16, dev)
        ops.nchw_to_nhwc(dout.contiguous().float(), dy, B, bp.C, H, W)
        grads = {p: torch.zeros_like(p) for p in ctx.params}
        dx = engine._e((M, Ci), torch.bfloat16, dev)
        engine.block_backward(bp, ctx.pk, ctx.bctx, ctx.xb, dy, None, dx, grads)
        dxo = engine._e((B, Ci, H, W), torch.float32, dev)
        ops.nhwc_to_nchw(dx, dxo, B, Ci, H, W)
        return (dxo, None, None) + tuple(grads[p] for p in ctx.params)


class DynamicFusionConvAttnBlock(nn.Module):
    """reference models/unet_dfc_sa_res.py:41-116."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, pool_size=8, ablation_on_qk_channels=8,
                 full_res_attention=False):
        super().__init__()
        if (kernel_size, stride, padding) != (3, 1, 1):
            raise NotImplementedError("dfcsa: the DFC-SA block kernels implement the 3x3 / stride 1 / pad 1 conv branch the "
                                      "reference network uses (models/unet_dfc_sa_res.py:132-156)")
        self.conv_branch = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding),
            nn.BatchNorm2d(out_channels), nn.ReLU(inplace=True))
        self.attn_branch = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=1), nn.BatchNorm2d(out_channels), nn.ReLU(inplace=True),
            FullResolutionAttention(out_channels) if full_res_attention else
            LightSelfAttention(out_channels, pool_size=pool_size, ablation_on_qk_channels=ablation_on_qk_channels))
        self.gate = nn.Sequential(nn.Conv2d(out_channels * 2, out_channels, kernel_size=1), nn.BatchNorm2d(out_channels), nn.Sigmoid())
        self.fusion_conv = nn.Sequential(nn.Conv2d(out_channels * 3, out_channels, kernel_size=1), nn.BatchNorm2d(out_channels),
                                         nn.ReLU(inplace=True))
        if in_channels != out_channels:
            self.residual_conv = nn.Conv2d(in_channels, out_channels, kernel_size=1, bias=False)
        else:
            self.residual_conv = nn.Identity()
        self.res_scale = nn.Parameter(torch.tensor(0.1))

    def forward(self, x):
        params = list(self.parameters())
        need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
        return _BlockFunction.apply(x, self, need_grad, *params)


class _NetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, net, need_grad, *params):
        if not net.training and not need_grad and getattr(net, "eval_cuda_graph", False) and x.is_cuda:
            logits, nctx = engine.eval_forward_graphed(net, x), None      # inference replayed from a CUDA graph (opt-in)
        else:
            logits, nctx = engine.net_forward(net, x, net.training, save=need_grad)
        ctx.net, ctx.nctx, ctx.params = net, nctx, params
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        if ctx.nctx is None:
            raise RuntimeError("dfcsa: backward through a network that ran in eval / no-grad mode")
        # one zeroed flat buffer carved into per-parameter views (128-byte aligned, the layout FusedSGD uses): one memset
        # instead of a fill kernel per parameter tensor (235 of them)
        from .ddp import flat_offsets
        offs, total = flat_offsets(list(ctx.params))
        flat = torch.zeros(total, dtype=torch.float32, device=dlogits.device)
        grads = {p: flat[offs[p][0]:offs[p][1]].view(p.shape) for p in ctx.params}
        engine.net_backward(ctx.net, ctx.nctx, dlogits, grads)
        ctx.nctx = None
        return (None, None, None) + tuple(grads[p] for p in ctx.params)


class UNetDFCSA(nn.Module):
    """reference models/unet_dfc_sa_res.py:118-204."""

    def __init__(self, in_channels=3, out_channels=1, features=[64, 128, 256, 512], pool_size=8, ablation_on_qk_channels=8,
                 full_res_attention=False):
        super().__init__()
        kw = dict(kernel_size=3, stride=1, padding=1, pool_size=pool_size, ablation_on_qk_channels=ablation_on_qk_channels,
                  full_res_attention=full_res_attention)
        self.down1 = DynamicFusionConvAttnBlock(in_channels, features[0], **kw)
        self.pool1 = nn.MaxPool2d(kernel_size=2, stride=2)
        self.down2 = DynamicFusionConvAttnBlock(features[0], features[1], **kw)
        self.pool2 = nn.MaxPool2d(kernel_size=2, stride=2)
        self.down3 = DynamicFusionConvAttnBlock(features[1], features[2], **kw)
        self.pool3 = nn.MaxPool2d(kernel_size=2, stride=2)
        self.down4 = DynamicFusionConvAttnBlock(features[2], features[3], **kw)
        self.pool4 = nn.MaxPool2d(kernel_size=2, stride=2)
        self.bottleneck = DynamicFusionConvAttnBlock(features[3], features[3] * 2, **kw)
        self.up4 = nn.ConvTranspose2d(features[3] * 2, features[3], kernel_size=2, stride=2)
        self.up_conv4 = DynamicFusionConvAttnBlock(features[3] * 2, features[3], **kw)
        self.up3 = nn.ConvTranspose2d(features[3], features[2], kernel_size=2, stride=2)
        self.up_conv3 = DynamicFusionConvAttnBlock(features[2] * 2, features[2], **kw)
        self.up2 = nn.ConvTranspose2d(features[2], features[1], kernel_size=2, stride=2)
        self.up_conv2 = DynamicFusionConvAttnBlock(features[1] * 2, features[1], **kw)
        self.up1 = nn.ConvTranspose2d(features[1], features[0], kernel_size=2, stride=2)
        self.up_conv1 = DynamicFusionConvAttnBlock(features[0] * 2, features[0], **kw)
        self.final_conv = nn.Conv2d(features[0], out_channels, kernel_size=1)

    def forward(self, x):
        params = list(self.parameters())
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _NetFunction.apply(x, self, need_grad, *params)


class UNetDFCSARes(UNetDFCSA):
    """reference models/unet_dfc_sa_res.py:207-220 (adds nothing to UNetDFCSA)."""


class FullResAttnDFCBlock(DynamicFusionConvAttnBlock):
    """reference models/unet_dfc_sa_ablation_attention.py:29-93: the DFC-SA block with full-resolution attention."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, **kwargs):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, full_res_attention=True)


class UNet_FullResAttention(UNetDFCSA):
    """reference models/unet_dfc_sa_ablation_attention.py:96-98 over AblationUNetBase
    (models/unet_dfc_sa_ablation_branches.py:104-164): same wiring and the same state_dict keys as UNetDFCSARes."""

    def __init__(self, in_channels, out_channels, features, **kwargs):
        super().__init__(in_channels, out_channels, features, full_res_attention=True)


class LocalOnlyBlock(nn.Module):
    """reference models/unet_dfc_sa_ablation_branches.py:72-101: relu(bn(conv3x3(x))) + res_scale * residual(x)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, **kwargs):
        super().__init__()
        if (kernel_size, stride, padding) != (3, 1, 1):
            raise NotImplementedError("dfcsa: LocalOnlyBlock kernels implement the 3x3 / stride 1 / pad 1 convolution")
        self.conv_branch = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding),
            nn.BatchNorm2d(out_channels), nn.ReLU(inplace=True))
        if in_channels != out_channels:
            self.residual_conv = nn.Conv2d(in_channels, out_channels, kernel_size=1, bias=False)
        else:
            self.residual_conv = nn.Identity()
        self.res_scale = nn.Parameter(torch.tensor(0.1))

    def forward(self, x):
        params = list(self.parameters())
        need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
        return _BlockFunction.apply(x, self, need_grad, *params)


class _AblationBlock(nn.Module):
    def forward(self, x):
        params = list(self.parameters())
        need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
        return _BlockFunction.apply(x, self, need_grad, *params)

    @staticmethod
    def _conv_branch(i, o):
        return nn.Sequential(nn.Conv2d(i, o, kernel_size=3, stride=1, padding=1), nn.BatchNorm2d(o), nn.ReLU(inplace=True))

    @staticmethod
    def _attn_branch(i, o, pool_size):
        return nn.Sequential(nn.Conv2d(i, o, kernel_size=1), nn.BatchNorm2d(o), nn.ReLU(inplace=True),
                             LightSelfAttention(o, pool_size=pool_size))

    def _residual(self, i, o):
        self.residual_conv = nn.Conv2d(i, o, kernel_size=1, bias=False) if i != o else nn.Identity()
        self.res_scale = nn.Parameter(torch.tensor(0.1))


class AttentionOnlyBlock(_AblationBlock):
    """reference models/unet_dfc_sa_ablation_branches.py:42-69: the attention branch + res_scale * residual."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, pool_size=8):
        super().__init__()
        self.attn_branch = self._attn_branch(in_channels, out_channels, pool_size)
        self._residual(in_channels, out_channels)


class AdditionFusionBlock(_AblationBlock):
    """reference models/unet_dfc_sa_ablation_fusion.py:9-55: conv branch + attention branch + res_scale * residual."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, pool_size=8):
        super().__init__()
        self.conv_branch = self._conv_branch(in_channels, out_channels)
        self.attn_branch = self._attn_branch(in_channels, out_channels, pool_size)
        self._residual(in_channels, out_channels)


class ConcatFusionBlock(_AblationBlock):
    """reference models/unet_dfc_sa_ablation_fusion.py:58-108: 1x1 conv + BN + ReLU over [conv branch | attention branch]."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, pool_size=8):
        super().__init__()
        self.conv_branch = self._conv_branch(in_channels, out_channels)
        self.attn_branch = self._attn_branch(in_channels, out_channels, pool_size)
        self.fusion_conv = nn.Sequential(nn.Conv2d(out_channels * 2, out_channels, kernel_size=1), nn.BatchNorm2d(out_channels),
                                         nn.ReLU(inplace=True))
        self._residual(in_channels, out_channels)


class _AblationUNet(nn.Module):
    """The U-Net wiring shared by the ablation networks (reference models/unet_dfc_sa_ablation_branches.py:104-164 and the
    three classes of models/unet_dfc_sa_ablation_placement.py): same sub-module names as UNetDFCSA, one block constructor
    for the encoder + bottleneck and one for the decoder."""

    def __init__(self, enc_block, dec_block, in_channels, out_channels, features):
        super().__init__()
        self.down1 = enc_block(in_channels, features[0])
        self.pool1 = nn.MaxPool2d(2)
        self.down2 = enc_block(features[0], features[1])
        self.pool2 = nn.MaxPool2d(2)
        self.down3 = enc_block(features[1], features[2])
        self.pool3 = nn.MaxPool2d(2)
        self.down4 = enc_block(features[2], features[3])
        self.pool4 = nn.MaxPool2d(2)
        self.bottleneck = enc_block(features[3], features[3] * 2)
        self.up4 = nn.ConvTranspose2d(features[3] * 2, features[3], kernel_size=2, stride=2)
        self.up_conv4 = dec_block(features[3] * 2, features[3])
        self.up3 = nn.ConvTranspose2d(features[3], features[2], kernel_size=2, stride=2)
        self.up_conv3 = dec_block(features[2] * 2, features[2])
        self.up2 = nn.ConvTranspose2d(features[2], features[1], kernel_size=2, stride=2)
        self.up_conv2 = dec_block(features[1] * 2, features[1])
        self.up1 = nn.ConvTranspose2d(features[1], features[0], kernel_size=2, stride=2)
        self.up_conv1 = dec_block(features[0] * 2, features[0])
        self.final_conv = nn.Conv2d(features[0], out_channels, kernel_size=1)

    def forward(self, x):
        params = list(self.parameters())
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _NetFunction.apply(x, self, need_grad, *params)


def _dfc(pool_size):
    return lambda i, o: DynamicFusionConvAttnBlock(i, o, pool_size=pool_size)


def _local(i, o):
    return LocalOnlyBlock(i, o)


class UNet_Baseline(_AblationUNet):
    """ablation 1(b), reference models/unet_dfc_sa_ablation_branches.py:166-168: LocalOnlyBlock everywhere."""

    def __init__(self, in_channels, out_channels, features, **kwargs):
        super().__init__(_local, _local, in_channels, out_channels, features)


class UNet_BothStandardConv(_AblationUNet):
    """ablation 4, reference models/unet_dfc_sa_ablation_placement.py:215-282: LocalOnlyBlock in encoder and decoder."""

    def __init__(self, in_channels, out_channels, features, **kwargs):
        super().__init__(_local, _local, in_channels, out_channels, features)


class UNet_EncoderOnlyDFC(_AblationUNet):
    """ablation 4, reference models/unet_dfc_sa_ablation_placement.py:83-147: DFC-SA blocks in the encoder + bottleneck,
    LocalOnlyBlock in the decoder."""

    def __init__(self, in_channels, out_channels, features, pool_size=8):
        super().__init__(_dfc(pool_size), _local, in_channels, out_channels, features)


class UNet_DecoderOnlyDFC(_AblationUNet):
    """ablation 4, reference models/unet_dfc_sa_ablation_placement.py:149-213: LocalOnlyBlock in the encoder + bottleneck,
    DFC-SA blocks in the decoder."""

    def __init__(self, in_channels, out_channels, features, pool_size=8):
        super().__init__(_local, _dfc(pool_size), in_channels, out_channels, features)


class UNet_AttentionOnly(_AblationUNet):
    """ablation 1(a), reference models/unet_dfc_sa_ablation_branches.py:170-171."""

    def __init__(self, in_channels, out_channels, features, pool_size=8):
        blk = lambda i, o: AttentionOnlyBlock(i, o, pool_size=pool_size)
        super().__init__(blk, blk, in_channels, out_channels, features)


class UNet_AdditionFusion(_AblationUNet):
    """ablation 2, reference models/unet_dfc_sa_ablation_fusion.py:102-104."""

    def __init__(self, in_channels, out_channels, features, pool_size=8):
        blk = lambda i, o: AdditionFusionBlock(i, o, pool_size=pool_size)
        super().__init__(blk, blk, in_channels, out_channels, features)


class UNet_ConcatFusion(_AblationUNet):
    """ablation 2, reference models/unet_dfc_sa_ablation_fusion.py:106-108."""

    def __init__(self, in_channels, out_channels, features, pool_size=8):
        blk = lambda i, o: ConcatFusionBlock(i, o, pool_size=pool_size)
        super().__init__(blk, blk, in_channels, out_channels, features)
