"""Kernel-level parity of the bandwidth-bound ("streaming") kernels of the DFC-SA block against plain torch fp32 on the
same 16-bit inputs, through the C ABI.  Shapes include the regimes of BASELINE.json's C2 config that the net-level tests
only reach through the whole network: pooled map LARGER than the feature map (s=14 -> P=16/32, s=28 -> P=32, reference
models/unet_dfc_sa_res.py:24,36), non-divisible windows (s=7 -> P=3, s=14 -> P=4), odd sizes and the level-1 224^2 map.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# (B, H, W, C, P)
SP = [(2, 14, 14, 64, 4), (2, 14, 14, 128, 16), (1, 14, 14, 64, 32), (1, 28, 28, 64, 32), (2, 7, 7, 64, 3),
      (1, 224, 224, 64, 4), (1, 9, 13, 40, 4), (1, 7, 5, 12, 8),
      # coarse pooled maps (H / P, W / P >= 8): the one-cell bilinear^T rows kernel and the CTA-per-cell column reductions,
      # with 16 / 32 channel vectors, a row count that leaves half a CTA idle, P = 8, and P = 3 (falls back to the two-sided rows)
      (1, 33, 72, 128, 4), (1, 72, 64, 256, 8), (3, 40, 48, 64, 3)]


def _rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-20))


def _nchw(x2d, B, H, W):
    return x2d.float().view(B, H, W, -1).permute(0, 3, 1, 2)


def _nhwc2d(x):
    return x.permute(0, 2, 3, 1).reshape(-1, x.shape[1])


def _affine(C, g, dev):
    return (torch.rand(C, generator=g) + 0.5).to(dev), (0.3 * torch.randn(C, generator=g)).to(dev)


def _bn_stats(x, g, dev):
    """(scale, shift, mean, invstd) of a train-mode BatchNorm over the rows of x with random affine parameters."""
    C = x.shape[1]
    mean = x.float().mean(0)
    invstd = (x.float().var(0, unbiased=False) + 1e-5).rsqrt()
    w, b = _affine(C, g, dev)
    return (w * invstd).contiguous(), (b - mean * w * invstd).contiguous(), mean.contiguous(), invstd.contiguous()


@pytest.mark.parametrize("B,H,W,C,P", SP)
def test_bnrelu_pool_fwd(cuda, B, H, W, C, P):
    """dfcsa_bnrelu_pool_fwd = adaptive_avg_pool2d(relu(bn2(A0)), P)  (reference :24 after :66-68)."""
    from dfcsa import ops
    g = torch.Generator().manual_seed(101)
    a0 = torch.randn(B * H * W, C, generator=g).to(cuda).half()
    sc, sh = _affine(C, g, cuda)
    tmp = torch.empty(B, H, P, C, device=cuda)
    pooled = torch.empty(B * P * P, C, device=cuda)
    ops.bnrelu_pool_fwd(a0, B, H, W, sc, sh, P, tmp, pooled)
    torch.cuda.synchronize()
    a = torch.relu(_nchw(a0, B, H, W) * sc.view(1, C, 1, 1) + sh.view(1, C, 1, 1))
    ref = _nhwc2d(F.adaptive_avg_pool2d(a, P))
    assert _rel(pooled, ref) < 1e-5
    assert (pooled - ref).abs().max().item() < 1e-4


@pytest.mark.parametrize("B,H,W,C,P", SP)
def test_branch_act_and_gate_mix_fwd(cuda, B, H, W, C, P):
    """dfcsa_branch_act_fwd: L = relu(bn1 L0), A = gamma * bilinear_up(o) + relu(bn2 A0) (reference :97,:99,:36,:38);
    dfcsa_gate_mix_fwd: f = g L + (1 - g) A with g = sigmoid(bn3 G0) (:104,:106).  z = [f | L | A], fp16."""
    from dfcsa import ops
    g = torch.Generator().manual_seed(102)
    M = B * H * W
    l0 = torch.randn(M, C, generator=g).to(cuda).half()
    a0 = torch.randn(M, C, generator=g).to(cuda).half()
    g0 = torch.randn(M, C, generator=g).to(cuda).half()
    o = torch.randn(B, P, P, C, generator=g).to(cuda)
    (s1, t1), (s2, t2), (s3, t3) = _affine(C, g, cuda), _affine(C, g, cuda), _affine(C, g, cuda)
    gamma = torch.tensor([0.6], device=cuda)
    z = torch.zeros(M, 3 * C, device=cuda, dtype=torch.float16)
    ops.branch_act_fwd(l0, a0, B, H, W, s1, t1, s2, t2, o, P, gamma, z, None)
    torch.cuda.synchronize()
    L = torch.relu(l0.float() * s1 + t1)
    U = _nhwc2d(F.interpolate(o.permute(0, 3, 1, 2), size=(H, W), mode="bilinear", align_corners=False))
    A = 0.6 * U + torch.relu(a0.float() * s2 + t2)
    assert (z[:, C:2 * C].float() - L).abs().max().item() <= 2e-3 * max(1.0, L.abs().max().item())
    assert (z[:, 2 * C:].float() - A).abs().max().item() <= 2e-3 * max(1.0, A.abs().max().item())
    assert _rel(z[:, C:2 * C], L) < 5e-4 and _rel(z[:, 2 * C:], A) < 5e-4
    ops.gate_mix_fwd(g0, s3, t3, z, None)
    torch.cuda.synchronize()
    gg = torch.sigmoid(g0.float() * s3 + t3)
    Ls, As = z[:, C:2 * C].float(), z[:, 2 * C:].float()          # the kernel mixes the stored fp16 L / A
    f = gg * Ls + (1 - gg) * As
    assert _rel(z[:, :C], f) < 6e-4
    assert (z[:, :C].float() - f).abs().max().item() <= 3e-3 * max(1.0, f.abs().max().item())


@pytest.mark.parametrize("pool", [False, True])
@pytest.mark.parametrize("B,H,W,C", [(2, 14, 14, 64), (1, 28, 28, 128), (1, 224, 224, 64), (2, 7, 9, 40), (1, 6, 4, 12), (1, 56, 56, 256)])
def test_block_out_and_sum_out_fwd(cuda, B, H, W, C, pool):
    """dfcsa_block_out_fwd: y = relu(bn4 F0) + res_scale * R with the fused 2x2 max pool (reference :110,:114,:164);
    dfcsa_sum_out_fwd: y = a [+ b] + res_scale * r (+ pool), the output stage of the sum ablation blocks."""
    from dfcsa import ops
    g = torch.Generator().manual_seed(103)
    M = B * H * W
    f0 = torch.randn(M, C, generator=g).to(cuda).half()
    r = torch.randn(M, C, generator=g).to(cuda).half()
    b2 = torch.randn(M, C, generator=g).to(cuda).half()
    s4, t4 = _affine(C, g, cuda)
    rs = torch.tensor(0.37, device=cuda)
    Hp, Wp = H // 2, W // 2
    for which in ("block", "sum2", "sum1"):
        y = torch.zeros(M, C, device=cuda, dtype=torch.float16)
        yp = torch.zeros(B * Hp * Wp, C, device=cuda, dtype=torch.float16) if pool else None
        if which == "block":
            ops.block_out_fwd(f0, r, B, H, W, s4, t4, rs, y, yp, None, None)
            ref = torch.relu(f0.float() * s4 + t4) + 0.37 * r.float()
        elif which == "sum2":
            ops.sum_out_fwd(f0, b2, r, B, H, W, rs, y, yp)
            ref = f0.float() + b2.float() + 0.37 * r.float()
        else:
            ops.sum_out_fwd(f0, None, r, B, H, W, rs, y, yp)
            ref = f0.float() + 0.37 * r.float()
        torch.cuda.synchronize()
        assert _rel(y, ref) < 5e-4, which
        assert (y.float() - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item()), which
        if pool:      # the pool is taken over the values as stored (fp16): exact against max_pool2d of the stored y
            want = _nhwc2d(F.max_pool2d(_nchw(y, B, H, W), 2))
            assert torch.equal(yp.float(), want), which


@pytest.mark.parametrize("src", ["skip", "pool", "both"])
@pytest.mark.parametrize("B,H,W,C", [(2, 14, 14, 64), (1, 28, 28, 128), (1, 112, 112, 64), (2, 8, 10, 40), (1, 56, 56, 256),
                                     (1, 9, 13, 64), (2, 7, 5, 16), (1, 3, 2, 8)])
def test_block_out_bwd_reduce_and_bn_bwd_apply(cuda, B, H, W, C, src):
    """dfcsa_block_out_bwd_reduce: dy = dskip + maxpool2x2^T(dyp) (argmax recomputed from the stored y, first maximum in
    scan order like ATen), BN4 reductions over d4 = dy [bn4(F0) > 0], d res_scale = sum dy R; then dfcsa_bn_bwd_apply:
    dF0 = scale (d4 - mean(d4) - xhat mean(d4 xhat)) - against autograd of y = relu(bn(F0)) + rs R, max_pool2d(y)."""
    from dfcsa import ops
    g = torch.Generator().manual_seed(104)
    M, Hp, Wp = B * H * W, H // 2, W // 2
    f0 = torch.randn(M, C, generator=g).to(cuda).half()
    r = torch.randn(M, C, generator=g).to(cuda).half()
    bn4 = _bn_stats(f0, g, cuda)
    rs = torch.tensor(0.37, device=cuda)
    y = torch.empty(M, C, device=cuda, dtype=torch.float16)
    yp = torch.empty(B * Hp * Wp, C, device=cuda, dtype=torch.float16)
    ops.block_out_fwd(f0, r, B, H, W, bn4[0], bn4[1], rs, y, yp, None, None)
    dskip = torch.randn(M, C, generator=g).to(cuda).bfloat16() if src != "pool" else None
    dyp = torch.randn(B * Hp * Wp, C, generator=g).to(cuda).bfloat16() if src != "skip" else None
    dy = dskip.clone() if dskip is not None else torch.zeros(M, C, device=cuda, dtype=torch.bfloat16)
    red = torch.zeros(2 * C + 1, device=cuda, dtype=torch.float64)
    ops.block_out_bwd_reduce(dy if dskip is not None else None, dyp, y, f0, r, B, H, W, *bn4, dy, red[:2 * C], red[2 * C:])
    dF0 = torch.empty(M, C, device=cuda, dtype=torch.bfloat16)
    ops.bn_bwd_apply(dy, f0, *bn4, red[:2 * C], 0, dF0)
    torch.cuda.synchronize()
    # reference
    dy_ref = dskip.float().clone() if dskip is not None else torch.zeros(M, C, device=cuda)
    if dyp is not None:
        yn = _nchw(y, B, H, W).clone().requires_grad_(True)
        (F.max_pool2d(yn, 2) * _nchw(dyp, B, Hp, Wp)).sum().backward()
        dy_ref = dy_ref + _nhwc2d(yn.grad)
    assert _rel(dy, dy_ref) < 4e-3
    dyr = dy.float()                                   # the later passes read the stored (bf16) dy
    sc, sh, mean, invstd = bn4
    d4 = torch.where(f0.float() * sc + sh > 0, dyr, torch.zeros_like(dyr))
    xhat = (f0.float() - mean) * invstd
    assert _rel(red[:C], d4.sum(0)) < 1e-3 and _rel(red[C:2 * C], (d4 * xhat).sum(0)) < 1e-3
    assert abs(float(red[2 * C]) - float((dyr * r.float()).sum())) <= 1e-3 * float((dyr * r.float()).abs().sum()) + 1e-3
    dF0_ref = sc * (d4 - d4.mean(0) - xhat * (d4 * xhat).mean(0))
    assert _rel(dF0, dF0_ref) < 6e-3


@pytest.mark.parametrize("M,C", [(2 * 14 * 14, 64), (28 * 28, 128), (112 * 112, 64), (2 * 9 * 7, 40), (56 * 56, 256)])
def test_gate_mix_bwd(cuda, M, C):
    """dfcsa_gate_mix_bwd_reduce / _apply: dS = df (L - A) g (1 - g), g = sigmoid(bn3 G0), BN3 backward of dS -> dG0
    (autograd of reference :104,:106 with a train-mode BatchNorm in front of the sigmoid)."""
    from dfcsa import ops
    g = torch.Generator().manual_seed(105)
    dz = torch.randn(M, 3 * C, generator=g).to(cuda).bfloat16()
    z = torch.randn(M, 3 * C, generator=g).to(cuda).half()
    g0 = torch.randn(M, C, generator=g).to(cuda).half()
    bn3 = _bn_stats(g0, g, cuda)
    red3 = torch.zeros(2 * C, device=cuda, dtype=torch.float64)
    ops.gate_mix_bwd_reduce(dz, z, g0, *bn3, red3)
    dg0 = torch.empty(M, C, device=cuda, dtype=torch.bfloat16)
    ops.gate_mix_bwd_apply(dz, z, g0, *bn3, red3, dg0)
    torch.cuda.synchronize()
    sc, sh, mean, invstd = bn3
    gg = torch.sigmoid(g0.float() * sc + sh)
    dS = dz[:, :C].float() * (z[:, C:2 * C].float() - z[:, 2 * C:].float()) * gg * (1 - gg)
    xhat = (g0.float() - mean) * invstd
    assert _rel(red3[:C], dS.sum(0)) < 1e-3 and _rel(red3[C:], (dS * xhat).sum(0)) < 2e-3
    ref = sc * (dS - dS.mean(0) - xhat * (dS * xhat).mean(0))
    assert _rel(dg0, ref) < 6e-3


@pytest.mark.parametrize("B,H,W,C,P", SP)
def test_branch_backward_at_every_pool_regime(cuda, B, H, W, C, P):
    """branch_bwd_reduce1 / reduce2 / apply with the gate terms (g0 given) at the (s, P) regimes above: d_o =
    gamma * bilinear^T(dA), d gamma, pool^T(dpooled) added into dA, both BatchNorm backward passes."""
    from dfcsa import ops
    dev = cuda
    g = torch.Generator().manual_seed(106)
    M = B * H * W
    dz = torch.randn(M, 3 * C, generator=g).to(dev).bfloat16()
    l0 = torch.randn(M, C, generator=g).to(dev).half()
    a0 = torch.randn(M, C, generator=g).to(dev).half()
    g0 = torch.randn(M, C, generator=g).to(dev).half()
    o = torch.randn(B, P, P, C, generator=g).to(dev)
    dpooled = torch.randn(B, P, P, C, generator=g).to(dev)
    gamma = torch.tensor([0.7], device=dev)
    bn1, bn2 = _bn_stats(l0, g, dev), _bn_stats(a0, g, dev)
    s3, t3 = _affine(C, g, dev)
    dz_in = dz.clone()
    red = torch.zeros(4 * C + 2, device=dev, dtype=torch.float64)
    red1, red2, dgam = red[:2 * C], red[2 * C:4 * C], red[4 * C + 1:4 * C + 2]
    tmp = torch.empty(B, H, P, C, device=dev)
    d_o = torch.empty(B * P * P, C, device=dev)
    ops.branch_bwd_reduce1(dz, l0, g0, B, H, W, *bn1, s3, t3, o, P, gamma, red1, dgam, tmp, d_o)
    ops.branch_bwd_reduce2(dz, a0, B, H, W, *bn2, dpooled, P, red2)
    dl0 = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    da0 = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    ops.branch_bwd_apply(dz, l0, a0, B, H, W, bn1, red1, bn2, red2, dpooled, P, dl0, da0)
    torch.cuda.synchronize()
    gg = torch.sigmoid(g0.float() * s3 + t3)
    df = dz_in[:, :C].float()
    dL_ref = dz_in[:, C:2 * C].float() + df * gg
    dA_ref = dz_in[:, 2 * C:].float() + df * (1 - gg)
    assert _rel(dz[:, C:2 * C], dL_ref) < 4e-3 and _rel(dz[:, 2 * C:], dA_ref) < 4e-3
    dL, dA = dz[:, C:2 * C].float(), dz[:, 2 * C:].float()      # as stored: what the later passes read
    o_n = o.permute(0, 3, 1, 2).clone().requires_grad_(True)
    U = F.interpolate(o_n, size=(H, W), mode="bilinear", align_corners=False)
    s = (U * _nchw(dA, B, H, W)).sum()
    s.backward()
    assert _rel(d_o, 0.7 * _nhwc2d(o_n.grad)) < 1e-4
    assert abs(float(dgam) - float(s)) <= 1e-4 * float((U * _nchw(dA, B, H, W)).abs().sum()) + 1e-3

    def bn_bwd(d, x, bnp):
        sc, sh, mean, invstd = bnp
        xf = x.float()
        dm = torch.where(xf * sc + sh > 0, d, torch.zeros_like(d))
        xhat = (xf - mean) * invstd
        return sc * (dm - dm.mean(0) - xhat * (dm * xhat).mean(0)), dm.sum(0), (dm * xhat).sum(0)

    dl0_ref, s1a, s1b = bn_bwd(dL, l0, bn1)
    assert _rel(red1[:C], s1a) < 1e-3 and _rel(red1[C:], s1b) < 2e-3
    assert _rel(dl0, dl0_ref) < 6e-3
    x_n = torch.zeros(B, C, H, W, device=dev, requires_grad=True)
    (F.adaptive_avg_pool2d(x_n, P) * dpooled.permute(0, 3, 1, 2)).sum().backward()
    da0_ref, s2a, s2b = bn_bwd(dA + _nhwc2d(x_n.grad), a0, bn2)
    assert _rel(red2[:C], s2a) < 1e-3 and _rel(red2[C:], s2b) < 2e-3
    assert _rel(da0, da0_ref) < 6e-3


@pytest.mark.parametrize("B,H,W,C,P", SP)
def test_window_terms_reduction_equals_gather_reduction(cuda, B, H, W, C, P):
    """The BatchNorm-2 backward sums without a gather pass: dfcsa_bn_bwd_reduce over (dA, A0) + dfcsa_pool_window_terms over
    the window means emitted by dfcsa_bnrelu_pool_fwd(with_masks) equal branch_bwd_reduce2's per-pixel pool^T gather (and the
    torch reference), including overlapping windows (s = 14, P = 4) and pooled maps larger than the feature map."""
    from dfcsa import ops
    g = torch.Generator().manual_seed(107)
    M = B * H * W
    dz = torch.randn(M, 3 * C, generator=g).to(cuda).bfloat16()
    a0 = torch.randn(M, C, generator=g).to(cuda).half()
    dpooled = torch.randn(B * P * P, C, generator=g).to(cuda)
    bn2 = _bn_stats(a0, g, cuda)
    tmp = torch.empty(3, B, H, P, C, device=cuda)
    pooled3 = torch.empty(3, B * P * P, C, device=cuda)
    ops.bnrelu_pool_fwd(a0, B, H, W, bn2[0], bn2[1], P, tmp, pooled3, with_masks=True)
    red_new = torch.zeros(2 * C, device=cuda, dtype=torch.float64)
    ops.bn_bwd_reduce(dz[:, 2 * C:], a0, *bn2, red_new)
    ops.pool_window_terms(dpooled, pooled3[1:], B, P, C, bn2[2], bn2[3], red_new)
    red_old = torch.zeros(2 * C, device=cuda, dtype=torch.float64)
    ops.branch_bwd_reduce2(dz, a0, B, H, W, *bn2, dpooled, P, red_old)
    torch.cuda.synchronize()
    # plane 0 is the ordinary pooled activation, planes 1 / 2 the pooled mask and mask * A0
    sc, sh, mean, invstd = bn2
    bnv = _nchw(a0, B, H, W) * sc.view(1, C, 1, 1) + sh.view(1, C, 1, 1)
    m = (bnv > 0).float()
    assert _rel(pooled3[0], _nhwc2d(F.adaptive_avg_pool2d(torch.relu(bnv), P))) < 1e-5
    assert _rel(pooled3[1], _nhwc2d(F.adaptive_avg_pool2d(m, P))) < 1e-5
    assert _rel(pooled3[2], _nhwc2d(F.adaptive_avg_pool2d(m * _nchw(a0, B, H, W), P))) < 1e-5
    # torch reference of the two sums
    x_n = torch.zeros(B, C, H, W, device=cuda, requires_grad=True)
    (F.adaptive_avg_pool2d(x_n, P) * dpooled.view(B, P, P, C).permute(0, 3, 1, 2)).sum().backward()
    d = (dz[:, 2 * C:].float() + _nhwc2d(x_n.grad)) * _nhwc2d(m)
    xhat = (a0.float() - mean) * invstd
    for red in (red_new, red_old):
        assert _rel(red[:C], d.sum(0)) < 1e-3 and _rel(red[C:], (d * xhat).sum(0)) < 2e-3
    assert _rel(red_new, red_old) < 1e-3
