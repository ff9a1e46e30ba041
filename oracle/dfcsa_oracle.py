"""ORACLE — test infrastructure only.  Never imported by the product path (dfc-sa-unet_b200/).

A CPU restatement, in plain functional PyTorch fp32 (or fp64), of the DFC-SA-Res-Block hot path of
YukiHataRin/DFC-SA-UNet.  Every function cites the reference lines it follows (paths relative to the reference
root).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
file, and only as the checker / the timed CPU baseline.

Pinning: the reference ships no tests or golden vectors, so parity is pinned by outputs of the
reference itself, generated in the build container by tests/golden/make_golden.py (imports /root/reference by file
path) and committed under tests/golden/*.npz; tests/test_oracle.py checks this restatement against them.

The adaptive-average-pool windows and the bilinear (align_corners=False) weights are restated with explicit index
arithmetic rather than by calling the ATen ops, because those index rules are exactly what the CUDA kernels have to
reproduce; tests/test_oracle.py checks them against ATen.
"""
import math

import torch
import torch.nn.functional as F

# bench.py's CPU-baseline leg sets this so the timed port issues the same ATen calls the reference issues
# (F.batch_norm, F.adaptive_avg_pool2d, F.interpolate) instead of the explicit restatements; tests check both agree.
USE_ATEN_OPS = False

# --------------------------------------------------------------------------------------------------------------
# index rules
# --------------------------------------------------------------------------------------------------------------


def adaptive_pool_matrix(s, P, dtype=torch.float32):
    """[P, s] averaging matrix of F.adaptive_avg_pool2d along one axis (reference models/unet_dfc_sa_res.py:24):
    window i = [floor(i*s/P), ceil((i+1)*s/P))."""
    m = torch.zeros(P, s, dtype=dtype)
    for i in range(P):
        lo = (i * s) // P
        hi = -((-(i + 1) * s) // P)
        m[i, lo:hi] = 1.0 / (hi - lo)
    return m


def bilinear_matrix(P, s, dtype=torch.float32):
    """[s, P] interpolation matrix of F.interpolate(mode='bilinear', align_corners=False) along one axis
    (reference models/unet_dfc_sa_res.py:36): src = max((d+0.5)*P/s - 0.5, 0); i0 = floor(src); i1 = min(i0+1, P-1)."""
    m = torch.zeros(s, P, dtype=dtype)
    scale = P / s
    for d in range(s):
        src = max((d + 0.5) * scale - 0.5, 0.0)
        i0 = min(int(math.floor(src)), P - 1)
        i1 = min(i0 + 1, P - 1)
        l1 = src - i0
        m[d, i0] += 1.0 - l1
        m[d, i1] += l1
    return m


def adaptive_avg_pool(x, P):
    if USE_ATEN_OPS:
        return F.adaptive_avg_pool2d(x, (P, P))
    B, C, H, W = x.shape
    my = adaptive_pool_matrix(H, P, x.dtype)
    mx = adaptive_pool_matrix(W, P, x.dtype)
    return torch.einsum("ph,bchw,qw->bcpq", my, x, mx)


def bilinear_upsample(o, H, W):
    if USE_ATEN_OPS:
        return F.interpolate(o, size=(H, W), mode="bilinear", align_corners=False)
    B, C, P, Q = o.shape
    my = bilinear_matrix(P, H, o.dtype)
    mx = bilinear_matrix(Q, W, o.dtype)
    return torch.einsum("hp,bcpq,wq->bchw", my, o, mx)


# --------------------------------------------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------------------------------------------


def batch_norm(x, sd, prefix, training, momentum=0.1, eps=1e-5, update_running=True):
    """nn.BatchNorm2d (reference models/unet_dfc_sa_res.py:60,67,75,82): batch statistics with biased variance for
    normalisation; running_var gets the unbiased variance; num_batches_tracked += 1."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    if USE_ATEN_OPS and (update_running or not training):
        if training:
            sd[prefix + ".num_batches_tracked"] += 1
        return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], w, b, training, momentum, eps)
    if training:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        if update_running:
            n = x.numel() / x.shape[1]
            with torch.no_grad():
                sd[prefix + ".running_mean"].mul_(1 - momentum).add_(momentum * mean.detach())
                sd[prefix + ".running_var"].mul_(1 - momentum).add_(momentum * var.detach() * n / max(n - 1, 1))
                sd[prefix + ".num_batches_tracked"] += 1
    else:
        mean, var = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    xhat = (x - mean[None, :, None, None]) * torch.rsqrt(var[None, :, None, None] + eps)
    return xhat * w[None, :, None, None] + b[None, :, None, None]


def light_self_attention(x, sd, prefix, pool_size):
    """LightSelfAttention.forward (reference models/unet_dfc_sa_res.py:20-39)."""
    B, C, H, W = x.shape
    pooled = adaptive_avg_pool(x, pool_size)                                          # :24
    N = pool_size * pool_size
    q = F.conv2d(pooled, sd[prefix + ".query_conv.weight"], sd[prefix + ".query_conv.bias"]).view(B, -1, N).permute(0, 2, 1)  # :28
    k = F.conv2d(pooled, sd[prefix + ".key_conv.weight"], sd[prefix + ".key_conv.bias"]).view(B, -1, N)                       # :29
    energy = torch.bmm(q, k)                                                          # :30 (no 1/sqrt(d))
    attention = torch.softmax(energy, dim=-1)                                         # :31
    v = F.conv2d(pooled, sd[prefix + ".value_conv.weight"], sd[prefix + ".value_conv.bias"]).view(B, C, N)                    # :32
    out = torch.bmm(v, attention.permute(0, 2, 1)).view(B, C, pool_size, pool_size)   # :33-34
    out = bilinear_upsample(out, H, W)                                                # :36
    return sd[prefix + ".gamma"] * out + x                                            # :38


def full_resolution_attention(x, sd, prefix):
    """FullResolutionAttention.forward (reference models/unet_dfc_sa_ablation_attention.py:15-26): the same
    attention without pooling / upsampling, N = H*W."""
    B, C, H, W = x.shape
    N = H * W
    q = F.conv2d(x, sd[prefix + ".query_conv.weight"], sd[prefix + ".query_conv.bias"]).view(B, -1, N).permute(0, 2, 1)
    k = F.conv2d(x, sd[prefix + ".key_conv.weight"], sd[prefix + ".key_conv.bias"]).view(B, -1, N)
    attention = torch.softmax(torch.bmm(q, k), dim=-1)
    v = F.conv2d(x, sd[prefix + ".value_conv.weight"], sd[prefix + ".value_conv.bias"]).view(B, C, N)
    out = torch.bmm(v, attention.permute(0, 2, 1)).view(B, C, H, W)
    return sd[prefix + ".gamma"] * out + x


def dfc_block(x, sd, prefix, pool_size, training=True, full_res_attention=False, update_running=True):
    """DynamicFusionConvAttnBlock.forward (reference models/unet_dfc_sa_res.py:95-116)."""
    p = prefix
    bn = lambda t, name: batch_norm(t, sd, p + name, training, update_running=update_running)
    local = F.relu(bn(F.conv2d(x, sd[p + ".conv_branch.0.weight"], sd[p + ".conv_branch.0.bias"], padding=1), ".conv_branch.1"))  # :97
    a = F.relu(bn(F.conv2d(x, sd[p + ".attn_branch.0.weight"], sd[p + ".attn_branch.0.bias"]), ".attn_branch.1"))
    if full_res_attention:
        attn = full_resolution_attention(a, sd, p + ".attn_branch.3")
    else:
        attn = light_self_attention(a, sd, p + ".attn_branch.3", pool_size)                                                     # :99
    combined = torch.cat([local, attn], dim=1)                                                                                 # :102
    gate = torch.sigmoid(bn(F.conv2d(combined, sd[p + ".gate.0.weight"], sd[p + ".gate.0.bias"]), ".gate.1"))                   # :104
    fused = gate * local + (1 - gate) * attn                                                                                   # :106
    fusion_in = torch.cat([fused, combined], dim=1)                                                                            # :109
    out = F.relu(bn(F.conv2d(fusion_in, sd[p + ".fusion_conv.0.weight"], sd[p + ".fusion_conv.0.bias"]), ".fusion_conv.1"))     # :110
    if (p + ".residual_conv.weight") in sd:
        res = F.conv2d(x, sd[p + ".residual_conv.weight"])                                                                     # :113
    else:
        res = x                                                                                                                # nn.Identity, :90
    return out + sd[p + ".res_scale"] * res                                                                                    # :114


def local_only_block(x, sd, prefix, training=True, update_running=True):
    """LocalOnlyBlock.forward (reference models/unet_dfc_sa_ablation_branches.py:92-101)."""
    p = prefix
    local = F.relu(batch_norm(F.conv2d(x, sd[p + ".conv_branch.0.weight"], sd[p + ".conv_branch.0.bias"], padding=1), sd,
                              p + ".conv_branch.1", training, update_running=update_running))
    res = F.conv2d(x, sd[p + ".residual_conv.weight"]) if (p + ".residual_conv.weight") in sd else x
    return local + sd[p + ".res_scale"] * res


def branch_ablation_block(x, sd, prefix, pool_size, training=True, update_running=True):
    """AttentionOnlyBlock (reference models/unet_dfc_sa_ablation_branches.py:60-69), AdditionFusionBlock
    (models/unet_dfc_sa_ablation_fusion.py:40-55) and ConcatFusionBlock (:88-100), told apart by their parameters."""
    p = prefix
    bn = lambda t, name: batch_norm(t, sd, p + name, training, update_running=update_running)
    a = F.relu(bn(F.conv2d(x, sd[p + ".attn_branch.0.weight"], sd[p + ".attn_branch.0.bias"]), ".attn_branch.1"))
    attn = light_self_attention(a, sd, p + ".attn_branch.3", pool_size)
    res = F.conv2d(x, sd[p + ".residual_conv.weight"]) if (p + ".residual_conv.weight") in sd else x
    if (p + ".conv_branch.0.weight") not in sd:
        return attn + sd[p + ".res_scale"] * res                                   # attention only
    local = F.relu(bn(F.conv2d(x, sd[p + ".conv_branch.0.weight"], sd[p + ".conv_branch.0.bias"], padding=1), ".conv_branch.1"))
    if (p + ".fusion_conv.0.weight") in sd:                                        # concat fusion
        fused = F.relu(bn(F.conv2d(torch.cat([local, attn], dim=1), sd[p + ".fusion_conv.0.weight"], sd[p + ".fusion_conv.0.bias"]),
                          ".fusion_conv.1"))
    else:                                                                          # addition fusion
        fused = local + attn
    return fused + sd[p + ".res_scale"] * res


def unet_forward(x, sd, pool_size, training=True, full_res_attention=False, update_running=True):
    """UNetDFCSA.forward (reference models/unet_dfc_sa_res.py:161-204).  Channel widths come from the tensors.
    The ablation networks share this wiring (models/unet_dfc_sa_ablation_branches.py:132-164,
    models/unet_dfc_sa_ablation_placement.py:108-147 etc.) and differ only in the block type per position, which is read
    off the state dict: a block without gate parameters is a LocalOnlyBlock."""
    def blk(t, name):
        if (name + ".gate.0.weight") in sd:
            return dfc_block(t, sd, name, pool_size, training, full_res_attention, update_running)
        if (name + ".attn_branch.0.weight") in sd:
            return branch_ablation_block(t, sd, name, pool_size, training, update_running)
        return local_only_block(t, sd, name, training, update_running)
    d1 = blk(x, "down1"); p1 = F.max_pool2d(d1, 2, 2)      # :163-164
    d2 = blk(p1, "down2"); p2 = F.max_pool2d(d2, 2, 2)
    d3 = blk(p2, "down3"); p3 = F.max_pool2d(d3, 2, 2)
    d4 = blk(p3, "down4"); p4 = F.max_pool2d(d4, 2, 2)
    u = blk(p4, "bottleneck")                              # :176
    for name, skip in (("4", d4), ("3", d3), ("2", d2), ("1", d1)):
        u = F.conv_transpose2d(u, sd[f"up{name}.weight"], sd[f"up{name}.bias"], stride=2)   # :179
        if u.shape[2:] != skip.shape[2:]:
            u = F.interpolate(u, size=skip.shape[2:], mode="bilinear", align_corners=False)  # :180-181
        u = blk(torch.cat([u, skip], dim=1), f"up_conv{name}")                              # :182-183
    return F.conv2d(u, sd["final_conv.weight"], sd["final_conv.bias"])                      # :203


# --------------------------------------------------------------------------------------------------------------
# loss / metrics / optimizer
# --------------------------------------------------------------------------------------------------------------


def dice_loss(pred, target, smooth=1.0):
    """reference utils/metrics.py:6-24 (whole batch flattened)."""
    pred = pred.reshape(-1)
    target = target.reshape(-1)
    inter = (pred * target).sum()
    return 1 - (2.0 * inter + smooth) / (pred.sum() + target.sum() + smooth)


def bce_loss(pred, target):
    """nn.BCELoss() as constructed at reference utils/metrics.py:60 and called at :74: mean reduction; ATen clamps the
    log terms at -100 in the forward and the denominator p*(1-p) at 1e-12 in the backward, so saturated probabilities
    give finite (zero-ish) gradients - restating it with torch.clamp(torch.log(p)) would give NaN gradients there."""
    return F.binary_cross_entropy(pred, target)


def calculate_metrics(pred, target, loss_type="bce_dice", loss_params=None):
    """reference utils/metrics.py:211-263, 'bce_dice' branch :245-249.  pred is the sigmoid output.  Note the
    reference reads 'weight_bce' / 'weight_dice' (the YAMLs spell them bce_weight / dice_weight, so both are 1.0)."""
    loss_params = loss_params or {}
    pred_binary = (pred > 0.5).to(pred.dtype)                                    # :228
    inter = (pred_binary * target).sum().item()                                  # :231
    union = (pred_binary + target).sum().item() - inter                          # :232
    iou = inter / (union + 1e-7)                                                 # :233
    dice = (2.0 * inter) / (pred_binary.sum().item() + target.sum().item() + 1e-7)  # :236
    if loss_type != "bce_dice":
        raise ValueError(f"oracle restates only the bce_dice loss, got {loss_type}")
    w_bce = loss_params.get("weight_bce", 1.0)                                   # :246
    w_dice = loss_params.get("weight_dice", 1.0)                                 # :247
    loss = w_bce * bce_loss(pred, target) + w_dice * dice_loss(pred, target)     # :74-78
    return {"loss": loss, "iou": iou, "dice": dice}


def clip_grad_norm_(grads, max_norm=1.0):
    """torch.nn.utils.clip_grad_norm_ as called at reference utils/trainer.py:149."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef)
    return total


def sgd_step(params, grads, bufs, lr=0.01, momentum=0.9, weight_decay=1e-4):
    """torch.optim.SGD as configured at reference train.py:73-78 (no nesterov, no dampening)."""
    with torch.no_grad():
        for i, (p, g) in enumerate(zip(params, grads)):
            d = g + weight_decay * p
            if bufs[i] is None:
                bufs[i] = d.clone()
            else:
                bufs[i].mul_(momentum).add_(d)
            p.add_(bufs[i], alpha=-lr)


PARAM_SUFFIXES = (".weight", ".bias", ".gamma", ".res_scale")


def param_names(sd):
    return [k for k in sd if k.endswith(PARAM_SUFFIXES)]


def train_step(sd, bufs, images, masks, pool_size, lr=0.01, momentum=0.9, weight_decay=1e-4, max_norm=1.0,
               loss_params=None, full_res_attention=False):
    """One iteration of Trainer.train_epoch (reference utils/trainer.py:120-151): forward, sigmoid,
    calculate_metrics, backward, clip_grad_norm_(1.0), SGD step.  `sd` is modified in place."""
    names = param_names(sd)
    params = [sd[k] for k in names]
    for p in params:
        p.requires_grad_(True)
        p.grad = None
    logits = unet_forward(images, sd, pool_size, True, full_res_attention)
    m = calculate_metrics(torch.sigmoid(logits), masks, "bce_dice", loss_params)
    grads = list(torch.autograd.grad(m["loss"], params, allow_unused=True))
    grads = [g if g is not None else torch.zeros_like(p) for g, p in zip(grads, params)]
    for p in params:
        p.requires_grad_(False)
    gnorm = clip_grad_norm_(grads, max_norm)
    if bufs is None:
        bufs = [None] * len(params)
    sgd_step(params, grads, bufs, lr, momentum, weight_decay)
    return {"loss": m["loss"].item(), "iou": m["iou"], "dice": m["dice"], "grad_norm": gnorm.item(),
            "logits": logits.detach(), "bufs": bufs, "names": names}


# --------------------------------------------------------------------------------------------------------------
# parameter initialisation (same tensor names / shapes / init distributions as the reference constructors; the
# bit-exact reference init is only needed for the golden fixtures, which store their weights)
# --------------------------------------------------------------------------------------------------------------


def init_state_dict(in_channels=3, out_channels=1, features=(64, 128, 256, 512), qk=8, seed=0, dtype=torch.float32):
    """state_dict with the reference's 343-entry layout (reference models/unet_dfc_sa_res.py:118-159)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, co, ci, k, bias=True):
        bound = 1.0 / math.sqrt(ci * k * k)
        sd[name + ".weight"] = (torch.rand(co, ci, k, k, generator=g, dtype=dtype) * 2 - 1) * bound
        if bias:
            sd[name + ".bias"] = (torch.rand(co, generator=g, dtype=dtype) * 2 - 1) * bound

    def bnorm(name, c):
        sd[name + ".weight"] = torch.ones(c, dtype=dtype)
        sd[name + ".bias"] = torch.zeros(c, dtype=dtype)
        sd[name + ".running_mean"] = torch.zeros(c, dtype=dtype)
        sd[name + ".running_var"] = torch.ones(c, dtype=dtype)
        sd[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    def block(name, ci, co):
        sd[name + ".res_scale"] = torch.tensor(0.1, dtype=dtype)
        conv(name + ".conv_branch.0", co, ci, 3); bnorm(name + ".conv_branch.1", co)
        conv(name + ".attn_branch.0", co, ci, 1); bnorm(name + ".attn_branch.1", co)
        sd[name + ".attn_branch.3.gamma"] = torch.zeros(1, dtype=dtype)
        conv(name + ".attn_branch.3.query_conv", co // qk, co, 1)
        conv(name + ".attn_branch.3.key_conv", co // qk, co, 1)
        conv(name + ".attn_branch.3.value_conv", co, co, 1)
        conv(name + ".gate.0", co, 2 * co, 1); bnorm(name + ".gate.1", co)
        conv(name + ".fusion_conv.0", co, 3 * co, 1); bnorm(name + ".fusion_conv.1", co)
        if ci != co:
            conv(name + ".residual_conv", co, ci, 1, bias=False)

    f = list(features)
    block("down1", in_channels, f[0]); block("down2", f[0], f[1]); block("down3", f[1], f[2]); block("down4", f[2], f[3])
    block("bottleneck", f[3], 2 * f[3])
    for name, ci, co in (("4", 2 * f[3], f[3]), ("3", f[3], f[2]), ("2", f[2], f[1]), ("1", f[1], f[0])):
        bound = 1.0 / math.sqrt(co * 4)
        sd[f"up{name}.weight"] = (torch.rand(ci, co, 2, 2, generator=g, dtype=dtype) * 2 - 1) * bound
        sd[f"up{name}.bias"] = (torch.rand(co, generator=g, dtype=dtype) * 2 - 1) * bound
        block(f"up_conv{name}", 2 * co, co)
    conv("final_conv", out_channels, f[0], 1)
    return sd


def synthetic_batch(B, H, W, seed=1, in_channels=3):
    """Structured synthetic images / masks: blob masks correlated with the image, so gradients
    are well conditioned."""
    g = torch.Generator().manual_seed(seed)
    low = F.interpolate(torch.randn(B, 1, 7, 7, generator=g), size=(H, W), mode="bicubic", align_corners=False)
    mask = (low > 0.3).float()
    image = 0.5 * torch.randn(B, in_channels, H, W, generator=g) + low
    return image, mask
