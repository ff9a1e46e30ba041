"""Kernel-level parity of the implicit-GEMM convolution, its weight gradient and the weight re-layout.

The tcgen05 path (backend TC) and the fp32 SIMT path are both checked against a plain PyTorch fp32 computation of
the same op on the same (already 16-bit-rounded) operands, so the only difference left is accumulation order:
tolerance 2e-3 relative to the output scale for the tensor-core path (fp32 accumulate, 16-bit output rounding).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel_err(a, b):
    a = a.float()
    b = b.float()
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-12)


def _pack_fwd(wt):  # [N, C, kh, kw] -> [N, (tap, c)]
    n, c, kh, kw = wt.shape
    return wt.permute(0, 2, 3, 1).reshape(n, kh * kw * c).contiguous()


def _run_conv(dev, B, H, W, cins, modes, N, backend, src_dt=torch.float16, w_dt=torch.float16, out_dt=torch.float16,
              bias=False, stats=False, accumulate=False, seed=0):
    from dfcsa import ops
    g = torch.Generator(device="cpu").manual_seed(seed)
    segs, ref = [], 0
    wparts = []
    for c, mode in zip(cins, modes):
        if mode == ops.TAP_2x2S2:
            x = torch.randn(B, 2 * H, 2 * W, c, generator=g).to(dev).to(src_dt)
            wt = (torch.randn(N, c, 2, 2, generator=g) / (4 * c) ** 0.5).to(dev).to(w_dt)
            ref = ref + F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), stride=2)
            segs.append((x.reshape(-1, c), mode))
        elif mode == ops.TAP_3x3:
            x = torch.randn(B, H, W, c, generator=g).to(dev).to(src_dt)
            wt = (torch.randn(N, c, 3, 3, generator=g) / (9 * c) ** 0.5).to(dev).to(w_dt)
            ref = ref + F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), padding=1)
            segs.append((x.reshape(-1, c), mode))
        else:
            x = torch.randn(B, H, W, c, generator=g).to(dev).to(src_dt)
            wt = (torch.randn(N, c, 1, 1, generator=g) / c ** 0.5).to(dev).to(w_dt)
            ref = ref + F.conv2d(x.float().permute(0, 3, 1, 2), wt.float())
            segs.append((x.reshape(-1, c), mode))
        wparts.append(_pack_fwd(wt))
    w = torch.cat(wparts, dim=1).contiguous()
    if backend == ops.BACKEND_SIMT:
        w = w.float()
    ref = ref.permute(0, 2, 3, 1).reshape(-1, N)  # [M, N]
    bias_t = None
    if bias:
        bias_t = torch.randn(N, generator=g).to(dev)
        ref = ref + bias_t
    out = torch.zeros(B * H * W, N, device=dev, dtype=out_dt)
    if accumulate:
        base = torch.randn(B * H * W, N, generator=g).to(dev).to(out_dt)
        out.copy_(base)
        ref = ref + base.float()
    st = torch.zeros(2 * N, device=dev, dtype=torch.float64) if stats else None
    ops.conv_gemm(B, H, W, segs, w, N, out, accumulate=accumulate, bias=bias_t, stats=st, backend=backend)
    torch.cuda.synchronize()
    err = _rel_err(out, ref)
    serr = 0.0
    if stats:
        s_ref = torch.cat([ref.double().sum(0), (ref.double() ** 2).sum(0)])
        serr = ((st - s_ref).abs() / (s_ref.abs() + 1e-3 * s_ref.abs().max())).max().item()
    return err, serr


CASES = [
    # B, H, W, cins, modes(str), N, extras
    (2, 9, 11, [64], "1", 64, {}),                       # flattened 1x1, M=198 (tail tile)
    (1, 16, 16, [128], "1", 128, {"stats": True}),
    (2, 12, 12, [192], "1", 64, {"stats": True}),        # fusion-conv shape at level 1
    (2, 8, 8, [64], "1", 192, {"out_dt": torch.bfloat16, "src_dt": torch.bfloat16, "w_dt": torch.bfloat16}),
    (2, 8, 8, [64], "1", 384, {}),                       # two n tiles of 192
    (1, 32, 32, [64], "3", 64, {"stats": True}),         # 3x3, 32x4 spatial tiles
    (2, 14, 14, [64], "3", 128, {"stats": True}),        # 14x14: partial boxes
    (1, 28, 28, [128], "3", 64, {}),
    (1, 56, 56, [64], "3", 64, {"stats": True}),
    (2, 10, 12, [64, 64, 128], "311", 64, {"out_dt": torch.bfloat16, "src_dt": torch.bfloat16, "w_dt": torch.bfloat16}),  # dgrad: 3 segments
    (2, 6, 10, [64], "2", 128, {"src_dt": torch.bfloat16, "w_dt": torch.bfloat16, "out_dt": torch.bfloat16}),  # ConvT dgrad gather
    (1, 16, 16, [128], "1", 64, {"bias": True, "out_dt": torch.float32}),
    (1, 16, 16, [64], "1", 128, {"accumulate": True, "out_dt": torch.bfloat16, "src_dt": torch.bfloat16, "w_dt": torch.bfloat16}),
    (1, 4, 4, [512], "3", 1024, {}),                     # deep K, several n tiles
    (3, 13, 21, [128], "3", 96, {"stats": True}),        # halo-box 3x3 path: ragged 8x16 patches, N not a power of two
    (2, 17, 9, [64, 64], "31", 128, {"stats": True}),    # 3x3 + 1x1 segments through the halo-box path, W < 16
    (1, 40, 24, [64], "3", 256, {"stats": True}),        # N = 256: the box-per-tap 3x3 path
    (2, 40, 40, [64], "1", 256, {"stats": True}),        # short-K statistics GEMM: 64-wide n tiles, one n tile per CTA
    (4, 64, 64, [128], "1", 128, {"stats": True}),       # same, more tiles than SMs (several pixel tiles per CTA)
    (1, 24, 24, [128], "1", 192, {"bias": True}),        # bias staged in shared memory (no statistics)
    # CTA pairs (cta_group::2, M = 256 per MMA): 256-wide n tiles with K >= 256
    (3, 17, 19, [256], "1", 512, {"stats": True}),       # flat 1x1, 8 pixel tiles, two n tiles, statistics flushed per tile
    (1, 30, 30, [128, 128], "11", 256, {}),              # two segments, one n tile
    (5, 9, 9, [64], "3", 256, {"stats": True}),          # 3x3 box per tap, ODD number of pixel tiles: the peer of the last pair is masked
    (2, 24, 24, [256], "3", 512, {"bias": True, "out_dt": torch.bfloat16, "src_dt": torch.bfloat16, "w_dt": torch.bfloat16}),
    (1, 64, 64, [512], "1", 1024, {"accumulate": True, "out_dt": torch.bfloat16, "src_dt": torch.bfloat16, "w_dt": torch.bfloat16}),
    (2, 33, 31, [64, 64, 128], "311", 256, {"out_dt": torch.bfloat16, "src_dt": torch.bfloat16, "w_dt": torch.bfloat16}),  # dgrad-like, ragged
    (8, 64, 64, [1024], "1", 256, {"stats": True}),      # 256 pixel tiles = 128 pair tiles on 74 pairs: several tiles per pair,
    (4, 112, 112, [1024], "1", 256, {}),                 # ... and 392 tiles: both TMEM accumulators are reused (needs the peer's remote arrivals)
    (4, 112, 112, [128, 128, 128], "311", 256, {"out_dt": torch.bfloat16, "src_dt": torch.bfloat16, "w_dt": torch.bfloat16}),  # level-2 dgrad at batch 4
    (4, 112, 112, [1024], "1", 256, {"stats": True}),    # wide 1x1 with statistics at batch 4
    (4, 56, 56, [256, 256, 256], "311", 512, {"out_dt": torch.bfloat16, "src_dt": torch.bfloat16, "w_dt": torch.bfloat16}),    # level-3 dgrad
    (4, 56, 56, [512], "3", 256, {"stats": True}),       # level-3 decoder 3x3
    (1, 151, 128, [1024], "1", 512, {"stats": True}),    # 151 pixel tiles (odd): the peer CTA of the last pair has no tile, two n tiles
    (5, 75, 75, [128], "3", 256, {"stats": True}),       # 3x3 spatial tiles smaller than 128 pixels, odd tile count (K = 1152)
    # resident weights (one n tile, >= 4 pixel tiles per SM, the weight tile fits beside three activation stages)
    (2, 224, 224, [128], "3", 64, {"stats": True}),      # level-1 decoder 3x3: 147 KB of weights, 3 halo stages
    (2, 200, 200, [64, 64], "31", 64, {"bias": True}),   # halo path, two segments (the k-block offsets of the second one)
    (2, 224, 224, [128], "1", 64, {"stats": True}),      # flat 1x1, two k blocks per stage
    (3, 180, 180, [192], "1", 64, {}),                   # K = 192: an odd number of k blocks with two per stage
    (2, 224, 224, [128], "1", 128, {"stats": True}),     # short-K statistics GEMM: one 64-wide n tile per CTA, each with its own resident half
    (2, 224, 224, [64, 64], "11", 128, {"accumulate": True, "out_dt": torch.bfloat16, "src_dt": torch.bfloat16, "w_dt": torch.bfloat16}),
    (2, 160, 160, [64], "3", 128, {"stats": True}),      # level-2 encoder 3x3 (N = 128, halo path)
    (2, 160, 160, [384], "1", 128, {"stats": True}),     # level-2 fusion conv: 98 KB of weights
]
_MODE = {"1": 0, "3": 1, "2": 2}


@pytest.mark.parametrize("backend", [1, 0], ids=["simt", "tc"])
@pytest.mark.parametrize("case", CASES, ids=[f"c{i}" for i in range(len(CASES))])
def test_conv_gemm(cuda, case, backend):
    B, H, W, cins, modes, N, kw = case
    err, serr = _run_conv(cuda, B, H, W, cins, [_MODE[m] for m in modes], N, backend, **kw)
    tol = 6e-3 if kw.get("out_dt") == torch.bfloat16 else 2e-3   # bf16 output rounding is 2^-9 relative
    assert err < tol, f"conv_gemm rel err {err}"
    assert serr < 2e-3, f"BN statistics rel err {serr}"


def test_conv_gemm_mixed_formats_rejected(cuda):
    """kind::f16 cannot mix fp16 and bf16 operands (measured: illegal instruction on B200), so the C ABI refuses the
    call instead of faulting; the SIMT path takes any mix."""
    with pytest.raises(RuntimeError, match="share one 16-bit format"):
        _run_conv(cuda, 1, 16, 16, [128], [0], 64, 0, src_dt=torch.float16, w_dt=torch.bfloat16, out_dt=torch.float32)
    err, _ = _run_conv(cuda, 1, 16, 16, [128], [0], 64, 1, src_dt=torch.float16, w_dt=torch.bfloat16, out_dt=torch.float32)
    assert err < 2e-3


@pytest.mark.parametrize("backend", [1, 0], ids=["simt", "tc"])
def test_convt_forward(cuda, backend):
    """ConvTranspose2d 2x2 stride 2 as one GEMM with a scatter epilogue (reference models/unet_dfc_sa_res.py:147)."""
    from dfcsa import ops
    dev = cuda
    B, H, W, Ci, Co = 2, 7, 9, 128, 64
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, H, W, Ci, generator=g).to(dev).half()
    wt = (torch.randn(Ci, Co, 2, 2, generator=g) / Ci ** 0.5).to(dev).half()
    bias = torch.randn(Co, generator=g).to(dev)
    ref = F.conv_transpose2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, stride=2).permute(0, 2, 3, 1).reshape(-1, Co)
    w = wt.permute(2, 3, 1, 0).reshape(4 * Co, Ci).contiguous()  # [(q, co), ci]
    if backend == 1:
        w = w.float()
    out = torch.zeros(B * 2 * H * 2 * W, Co, device=dev, dtype=torch.float16)
    ops.conv_gemm(B, H, W, [(x.reshape(-1, Ci), 0)], w, 4 * Co, out, out_mode=ops.OUT_CONVT2x2, bias=bias, backend=backend)
    torch.cuda.synchronize()
    assert _rel_err(out, ref) < 2e-3


WG_CASES = [
    # B, H, W, C, N, x_mode, dy_mode
    (2, 9, 11, 64, 64, 0, 0),
    (1, 16, 16, 128, 128, 0, 0),
    (2, 8, 8, 192, 64, 0, 0),
    (1, 8, 8, 64, 256, 0, 0),
    (1, 32, 32, 64, 64, 1, 0),
    (2, 14, 14, 64, 128, 1, 0),
    (1, 28, 28, 128, 64, 1, 0),
    (2, 6, 10, 128, 64, 0, 2),
    (1, 4, 4, 512, 128, 1, 0),
    # enough pixel blocks for the wave-aware split selection to choose several CTAs per (tap row, tile)
    (4, 64, 64, 128, 64, 1, 0),
    (1, 96, 96, 192, 128, 0, 0),
]


@pytest.mark.parametrize("backend", [1, 0], ids=["simt", "tc"])
@pytest.mark.parametrize("case", WG_CASES, ids=[f"w{i}" for i in range(len(WG_CASES))])
def test_conv_wgrad(cuda, case, backend):
    from dfcsa import ops
    dev = cuda
    B, H, W, Cc, N, xm, dm = case
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, H, W, Cc, generator=g).to(dev).bfloat16()
    alpha = torch.tensor([0.37], device=dev)
    if dm == 2:
        dy = torch.randn(B, 2 * H, 2 * W, N, generator=g).to(dev).bfloat16()
        # dW[ci, co, q] of conv_transpose2d == weight grad of the stride-2 conv dy -> x
        xr = x.float().permute(0, 3, 1, 2).requires_grad_(False)
        wt = torch.zeros(Cc, N, 2, 2, device=dev, requires_grad=True)
        y = F.conv_transpose2d(xr, wt, stride=2)
        (gw,) = torch.autograd.grad(y, wt, dy.float().permute(0, 3, 1, 2))
        ref = gw.permute(1, 2, 3, 0).reshape(N, 4 * Cc)  # [n, (q, c)]
        taps = 4
    else:
        dy = torch.randn(B, H, W, N, generator=g).to(dev).bfloat16()
        k = 3 if xm == 1 else 1
        wt = torch.zeros(N, Cc, k, k, device=dev, requires_grad=True)
        y = F.conv2d(x.float().permute(0, 3, 1, 2), wt, padding=k // 2)
        (gw,) = torch.autograd.grad(y, wt, dy.float().permute(0, 3, 1, 2))
        ref = gw.permute(0, 2, 3, 1).reshape(N, k * k * Cc)
        taps = k * k
    ref = ref * 0.37
    dw = torch.zeros(N, taps * Cc, device=dev)
    if backend == 0:
        ops.conv_wgrad(B, H, W, x.reshape(-1, Cc), xm, dy.reshape(-1, N), dm, dw, alpha=alpha, backend=0)
    else:
        ops.conv_wgrad(B, H, W, x.reshape(-1, Cc), xm, dy.reshape(-1, N), dm, dw, alpha=alpha, backend=1)
    torch.cuda.synchronize()
    assert _rel_err(dw, ref) < 2e-3


def test_permute3(cuda):
    from dfcsa import ops
    dev = cuda
    w = torch.randn(24, 40, 3, 3, device=dev)
    fwd = torch.empty(24, 9 * 40, device=dev, dtype=torch.float16)
    ops.permute3(w, fwd, (24, 9, 40), (40 * 9, 1, 9))
    assert torch.equal(fwd, w.permute(0, 2, 3, 1).reshape(24, 360).half())
    dg = torch.empty(40, 9 * 24, device=dev, dtype=torch.bfloat16)
    sc = torch.tensor([0.5], device=dev)
    ops.permute3(w, dg, (40, 9, 24), (9, 1, 40 * 9), flip1=True, scale=sc)
    ref = (0.5 * w.flip(2, 3)).permute(1, 2, 3, 0).reshape(40, 216).bfloat16()
    assert torch.equal(dg, ref)


SMALL_CASES = [
    # the bandwidth-shaped kernels behind the SIMT backend (csrc/small.cu): first layer (Ci=3) and final conv (Co=1)
    (2, 9, 11, [3], "3", 64, {"src_dt": torch.float32, "stats": True}),
    (1, 20, 33, [3], "1", 128, {"src_dt": torch.float32, "stats": True}),
    (2, 7, 9, [64], "1", 1, {"out_dt": torch.float32, "bias": True}),
    (2, 7, 9, [128], "1", 2, {"out_dt": torch.float32, "bias": True, "src_dt": torch.bfloat16}),
    (2, 7, 9, [1], "1", 64, {"src_dt": torch.bfloat16, "out_dt": torch.bfloat16}),
    (2, 40, 41, [64], "1", 1, {"out_dt": torch.float32, "bias": True}),      # several pixels in flight per thread, ragged tail
]


@pytest.mark.parametrize("case", SMALL_CASES, ids=[f"s{i}" for i in range(len(SMALL_CASES))])
def test_conv_small_channel_kernels(cuda, case):
    B, H, W, cins, modes, N, kw = case
    err, serr = _run_conv(cuda, B, H, W, cins, [_MODE[m] for m in modes], N, 1, **kw)
    tol = 6e-3 if kw.get("out_dt") == torch.bfloat16 else 2e-3
    assert err < tol, f"conv small rel err {err}"
    assert serr < 2e-3, f"BN statistics rel err {serr}"


@pytest.mark.parametrize("case", [(2, 9, 11, 3, 64, 1), (1, 20, 33, 3, 128, 0), (2, 7, 9, 64, 1, 0), (3, 50, 47, 3, 64, 1),
                                  (4, 64, 61, 3, 64, 0), (2, 48, 50, 64, 1, 0), (1, 30, 31, 128, 1, 0)],
                         ids=["x3_3x3", "x3_1x1", "dy1", "x3_3x3_big", "x3_1x1_big", "dy1_big", "dy1_c128"])
def test_wgrad_small_channel_kernels(cuda, case):
    from dfcsa import ops
    dev = cuda
    B, H, W, Cc, N, xm = case
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, H, W, Cc, generator=g).to(dev)
    x = x if Cc == 3 else x.bfloat16()
    dy = torch.randn(B, H, W, N, generator=g).to(dev).bfloat16()
    k = 3 if xm == 1 else 1
    wt = torch.zeros(N, Cc, k, k, device=dev, requires_grad=True)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), wt, padding=k // 2)
    (gw,) = torch.autograd.grad(y, wt, dy.float().permute(0, 3, 1, 2))
    ref = gw.permute(0, 2, 3, 1).reshape(N, k * k * Cc) * 0.5
    dw = torch.zeros(N, k * k * Cc, device=dev)
    ops.conv_wgrad(B, H, W, x.reshape(-1, Cc), xm, dy.reshape(-1, N), 0, dw, alpha=torch.tensor([0.5], device=dev), backend=1)
    torch.cuda.synchronize()
    assert _rel_err(dw, ref) < 2e-3


@pytest.mark.parametrize("backend", [1, 0], ids=["simt", "tc"])
@pytest.mark.parametrize("C,N1,N2,cb,M", [(192, 64, 64, 64, 1500), (384, 128, 128, 128, 700), (128, 64, 64, 0, 900), (768, 256, 256, 256, 300)])
def test_conv_wgrad_two_gradients_one_launch(cuda, C, N1, N2, cb, M, backend):
    """dw[n, c] += a1 * sum_m dy[m, n] x[m, c]  and  dw2[n, c - cb] += sum_m dy2[m, n] x[m, c] (c >= cb) in one call."""
    from dfcsa import ops
    g = torch.Generator().manual_seed(11)
    x = torch.randn(M, C, generator=g).cuda().bfloat16()
    dy = torch.randn(M, N1, generator=g).cuda().bfloat16()
    dy2 = torch.randn(M, N2, generator=g).cuda().bfloat16()
    alpha = torch.tensor([0.25], device=cuda)
    dw = torch.zeros(N1, C, device=cuda)
    dw2 = torch.zeros(N2, C - cb, device=cuda)
    ops.conv_wgrad(1, 1, M, x, 0, dy, 0, dw, alpha=alpha, backend=backend, second=(dy2, dw2, cb, None))
    torch.cuda.synchronize()
    assert _rel_err(dw, 0.25 * dy.float().t() @ x.float()) < 2e-3
    assert _rel_err(dw2, dy2.float().t() @ x.float()[:, cb:]) < 2e-3


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("batch,M,N,K", [(3, 200, 72, 8), (2, 256, 320, 256), (1, 130, 64, 200), (2, 64, 8, 64)])
def test_bgemm_all_layouts(cuda, batch, M, N, K, a_mn, b_mn, dtype):
    """dfcsa_bgemm: C[b] = A[b] @ B[b] with either operand stored K-major or MN-major, ragged M / K, tiny K and N."""
    from dfcsa import ops
    g = torch.Generator().manual_seed(13)
    A = torch.randn(batch, M, K, generator=g).cuda().to(dtype)
    Bm = torch.randn(batch, K, N, generator=g).cuda().to(dtype)
    ref = torch.bmm(A.float(), Bm.float())
    pad = lambda n: (n + 7) // 8 * 8       # pitches must be multiples of 8 elements
    if a_mn:       # stored [K, M]
        As = torch.zeros(batch, K, pad(M), device=cuda, dtype=dtype); As[:, :, :M] = A.transpose(1, 2)
        a_b, ld_a = K * pad(M), pad(M)
    else:          # stored [M, K]
        As = torch.zeros(batch, M, pad(K), device=cuda, dtype=dtype); As[:, :, :K] = A
        a_b, ld_a = M * pad(K), pad(K)
    if b_mn:       # stored [K, N]
        Bs = torch.zeros(batch, K, pad(N), device=cuda, dtype=dtype); Bs[:, :, :N] = Bm
        b_b, ld_b = K * pad(N), pad(N)
    else:          # stored [N, K]
        Bs = torch.zeros(batch, N, pad(K), device=cuda, dtype=dtype); Bs[:, :, :K] = Bm.transpose(1, 2)
        b_b, ld_b = N * pad(K), pad(K)
    for cdt, tol in ((torch.float32, 2e-3), (dtype, 1.2e-2)):
        Cm = torch.full((batch, M, N), float("nan"), device=cuda, dtype=cdt)
        ops.bgemm(batch, M, N, K, As, a_b, ld_a, a_mn, Bs, b_b, ld_b, b_mn, Cm, M * N, N)
        torch.cuda.synchronize()
        assert _rel_err(Cm, ref) < tol


@pytest.mark.parametrize("batch,N,K", [(2, 64, 8), (1, 200, 16), (3, 256, 8), (1, 1048, 64), (1, 3136, 8)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_softmax_bgemm(cuda, batch, N, K, dtype):
    """softmax(q k^T) through the ROWSTATS / EXP epilogues of dfcsa_bgemm equals torch.softmax of the fp32 product of
    the same 16-bit operands; rows sum to one; ragged tiles (N not a multiple of the tile) and multi-tile rows."""
    from dfcsa import ops
    g = torch.Generator().manual_seed(23)
    q = (torch.randn(batch, N, K, generator=g) * 1.5).cuda().to(dtype)
    k = (torch.randn(batch, N, K, generator=g) * 1.5).cuda().to(dtype)
    ref = torch.softmax(torch.bmm(q.float(), k.float().transpose(1, 2)), -1)
    for odt, tol in ((torch.float16, 2e-3), (torch.bfloat16, 8e-3)):
        out = torch.full((batch, N, N), float("nan"), device=cuda, dtype=odt)
        ops.softmax_bgemm(batch, N, N, K, q, N * K, K, k, N * K, K, out)
        torch.cuda.synchronize()
        assert _rel_err(out, ref) < tol
        assert (out.float().sum(-1) - 1).abs().max() < (4e-3 if odt == torch.float16 else 2e-2)


@pytest.mark.parametrize("batch,N,C", [(2, 64, 64), (1, 200, 128), (2, 1048, 64), (1, 3136, 64)])
@pytest.mark.parametrize("pdt", [torch.float16, torch.bfloat16])
def test_softmax_bwd_bgemm(cuda, batch, N, C, pdt):
    """dS = P * (dO V^T - D) from the SOFTMAX_BWD epilogue of dfcsa_bgemm, in place over P when the types allow."""
    from dfcsa import ops
    g = torch.Generator().manual_seed(29)
    P = torch.softmax(torch.randn(batch, N, N, generator=g) * 2, -1).cuda().to(pdt)
    dO = torch.randn(batch, N, C, generator=g).cuda().to(torch.bfloat16)
    V = torch.randn(batch, N, C, generator=g).cuda().to(torch.bfloat16)
    D = torch.randn(batch, N, generator=g).cuda()
    ref = P.float() * (torch.bmm(dO.float(), V.float().transpose(1, 2)) - D[..., None])
    dS = P.clone() if pdt == torch.bfloat16 else torch.full((batch, N, N), float("nan"), device=cuda, dtype=torch.bfloat16)
    ops.softmax_bwd_bgemm(batch, N, N, C, dO, N * C, C, V, N * C, C, dS if pdt == torch.bfloat16 else P, D.view(-1), dS)
    torch.cuda.synchronize()
    assert _rel_err(dS, ref) < 8e-3


@pytest.mark.parametrize("xdt,ydt", [(torch.float32, torch.float16), (torch.float32, torch.bfloat16), (torch.float16, torch.bfloat16),
                                     (torch.float16, torch.float32), (torch.bfloat16, torch.float32)])
@pytest.mark.parametrize("M,C,pad", [(37, 64, 0), (1000, 1024, 0), (50, 24, 8), (33, 20, 4)])
def test_cast2d(cuda, xdt, ydt, M, C, pad):
    """dfcsa_cast2d (vector and scalar paths, pitched rows): same values as torch's conversion."""
    from dfcsa import ops
    g = torch.Generator().manual_seed(31)
    xs = torch.randn(M, C + pad, generator=g).cuda().to(xdt)
    ys = torch.zeros(M, C + pad, device=cuda, dtype=ydt)
    ops.cast2d(xs[:, :C], ys[:, :C])
    assert torch.equal(ys[:, :C], xs[:, :C].to(ydt)) and (ys[:, C:] == 0).all()


@pytest.mark.parametrize("batch,N,Cq,C", [(1, 128, 8, 64), (2, 200, 8, 64), (1, 1048, 16, 128), (3, 64, 8, 64), (1, 3136, 8, 64),
                                           (1, 12544, 16, 128)])
def test_attn_pv_fused(cuda, batch, N, Cq, C):
    """dfcsa_attn_pv_fused: softmax(q k^T) v with the probabilities kept on chip, against torch on the same fp16 operands
    (ragged last tiles, several images, both channel widths)."""
    from dfcsa import ops
    g = torch.Generator().manual_seed(37)
    nq = 2 * Cq + C
    qkv = torch.randn(batch * N, nq, generator=g).cuda()
    qkv[:, :2 * Cq] *= 1.3
    q16 = qkv.half()
    q, k, v = (t.float().view(batch, N, -1) for t in (q16[:, :Cq], q16[:, Cq:2 * Cq], q16[:, 2 * Cq:]))
    ref = torch.bmm(torch.softmax(torch.bmm(q, k.transpose(1, 2)), -1), v)
    lse = torch.empty(batch * N, device=cuda)
    ops.attn_row_lse(batch, N, Cq, q16, nq, lse)
    assert torch.allclose(lse.view(batch, N), torch.logsumexp(torch.bmm(q, k.transpose(1, 2)), -1), atol=2e-3)
    o = torch.full((batch, N, C), float("nan"), device=cuda)
    ops.attn_pv_fused(q16, batch, N, Cq, C, lse, o)
    torch.cuda.synchronize()
    assert _rel_err(o, ref) < 3e-3


@pytest.mark.parametrize("batch,N,Cq,C", [(1, 128, 8, 64), (2, 200, 8, 64), (1, 1048, 16, 128), (3, 64, 8, 64), (1, 3136, 8, 64),
                                           (1, 6272, 16, 128)])
def test_attn_bwd_fused(cuda, batch, N, Cq, C):
    """dfcsa_attn_bwd_fused: dq / dk / dv of o = softmax(q k^T) v against torch autograd on the same fp16 operands
    (ragged last tiles in both directions, several images, both channel widths)."""
    from dfcsa import ops
    g = torch.Generator().manual_seed(41)
    nq = 2 * Cq + C
    qkv = torch.randn(batch * N, nq, generator=g).cuda()
    qkv[:, :2 * Cq] *= 1.3
    q16 = qkv.half()
    x = q16.float().view(batch, N, nq).clone().requires_grad_(True)
    q, k, v = x[..., :Cq], x[..., Cq:2 * Cq], x[..., 2 * Cq:]
    S = torch.bmm(q, k.transpose(1, 2))
    o = torch.bmm(torch.softmax(S, -1), v)
    dO = (torch.randn(batch, N, C, generator=g) * 0.05).cuda().bfloat16()
    o.backward(dO.float())
    ref = x.grad.view(batch * N, nq)
    lse = torch.logsumexp(S.detach(), -1).reshape(-1).contiguous()
    D = (dO.float() * o.detach()).sum(-1).reshape(-1).contiguous()
    dqkv = torch.full((batch * N, nq), float("nan"), device=cuda)
    ops.attn_bwd_fused(q16, q16.bfloat16(), dO.view(batch * N, C), batch, N, Cq, C, lse, D, dqkv)
    torch.cuda.synchronize()
    for name, lo, hi in (("dq", 0, Cq), ("dk", Cq, 2 * Cq), ("dv", 2 * Cq, nq)):
        assert _rel_err(dqkv[:, lo:hi], ref[:, lo:hi]) < 2e-2, name


@pytest.mark.parametrize("cols", [16, 100, 4096])
def test_softmax_rows_16bit(cuda, cols):
    from dfcsa import ops
    g = torch.Generator().manual_seed(17)
    x = (torch.randn(37, cols, generator=g) * 4).cuda()
    ref = torch.softmax(x, -1)
    for dt, tol in ((torch.float32, 1e-5), (torch.float16, 1e-3), (torch.bfloat16, 8e-3)):
        y = torch.empty(37, cols, device=cuda, dtype=dt)
        ops.softmax_rows(x, y)
        assert _rel_err(y, ref) < tol
        dy = torch.randn(37, cols, generator=g).cuda()
        dref = ref * (dy - (dy * ref).sum(-1, keepdim=True))
        for ddt, dtol in ((torch.float32, 2e-2 if dt != torch.float32 else 1e-5), (torch.bfloat16, 2e-2)):
            dx = torch.empty(37, cols, device=cuda, dtype=ddt)
            ops.softmax_rows_bwd(y, dy, dx)
            assert _rel_err(dx, dref) < dtol


@pytest.mark.parametrize("B,N,Cq,C", [(3, 16, 8, 64), (2, 16, 128, 1024), (2, 9, 4, 40), (1, 32, 16, 128)])
def test_attn_small_matches_torch(cuda, B, N, Cq, C):
    """dfcsa_attn_small_fwd / bwd (one kernel per direction for N <= 32) against autograd of softmax(q k^T) v."""
    from dfcsa import ops
    g = torch.Generator().manual_seed(21)
    ld = 2 * Cq + C
    qkv = torch.randn(B * N, ld, generator=g).cuda()
    d_o = torch.randn(B * N, C, generator=g).cuda()
    ref_in = qkv.clone().requires_grad_(True)
    r = ref_in.view(B, N, ld)
    q, k, v = r[..., :Cq], r[..., Cq:2 * Cq], r[..., 2 * Cq:]
    attn_ref = torch.softmax(q @ k.transpose(1, 2), -1)
    o_ref = attn_ref @ v
    o_ref.backward(d_o.view(B, N, C))
    attn = torch.empty(B, N, N, device=cuda)
    o = torch.empty(B * N, C, device=cuda)
    ops.attn_small_fwd(qkv, B, N, Cq, C, attn, o)
    dqkv = torch.empty_like(qkv)
    ops.attn_small_bwd(qkv, attn, d_o, B, N, Cq, C, dqkv)
    torch.cuda.synchronize()
    assert _rel_err(attn, attn_ref) < 1e-5 and _rel_err(o.view(B, N, C), o_ref) < 1e-5
    assert _rel_err(dqkv, ref_in.grad) < 1e-4
    # the optional bias-gradient outputs ACCUMULATE the column sums of dq / dk / dv (q / k / v conv biases)
    dbq, dbk, dbv = torch.full((Cq,), 0.5, device=cuda), torch.zeros(Cq, device=cuda), torch.zeros(C, device=cuda)
    dqkv2 = torch.empty_like(qkv)
    ops.attn_small_bwd(qkv, attn, d_o, B, N, Cq, C, dqkv2, dbq, dbk, dbv)
    torch.cuda.synchronize()
    gsum = ref_in.grad.sum(0)
    assert torch.equal(dqkv2, dqkv)
    assert _rel_err(dbq, gsum[:Cq] + 0.5) < 1e-4 and _rel_err(dbv, gsum[2 * Cq:]) < 1e-4
    # the key bias shifts every logit of a row by the same amount, so its gradient is identically zero (rows of dS sum to 0):
    # what is left is rounding noise, compared on the scale of the summands
    assert (dbk - gsum[Cq:2 * Cq]).abs().max().item() <= 1e-5 * dqkv[:, Cq:2 * Cq].abs().sum(0).max().item()


@pytest.mark.parametrize("case", [(1, 32, 32, 128, 64, 1, 0), (2, 9, 11, 256, 320, 0, 0), (2, 6, 10, 128, 64, 0, 2), (1, 20, 20, 64, 128, 1, 0)])
def test_conv_wgrad_fp16_activation_bf16_gradient(cuda, case):
    """x fp16 (a forward activation) with dy bf16: the kernel converts the x tiles to bf16 in shared memory, so the result
    equals the gradient computed from bf16-rounded x."""
    from dfcsa import ops
    B, H, W, Cc, N, xm, dm = case
    g = torch.Generator().manual_seed(23)
    x16 = torch.randn(B, H, W, Cc, generator=g).cuda().half()
    xr = x16.bfloat16().float()                       # what the tensor core sees
    if dm == 2:
        dy = torch.randn(B, 2 * H, 2 * W, N, generator=g).cuda().bfloat16()
        wt = torch.zeros(Cc, N, 2, 2, device=cuda, requires_grad=True)
        y = F.conv_transpose2d(xr.permute(0, 3, 1, 2), wt, stride=2)
        (gw,) = torch.autograd.grad(y, wt, dy.float().permute(0, 3, 1, 2))
        ref, taps = gw.permute(1, 2, 3, 0).reshape(N, 4 * Cc), 4
    else:
        dy = torch.randn(B, H, W, N, generator=g).cuda().bfloat16()
        k = 3 if xm == 1 else 1
        wt = torch.zeros(N, Cc, k, k, device=cuda, requires_grad=True)
        y = F.conv2d(xr.permute(0, 3, 1, 2), wt, padding=k // 2)
        (gw,) = torch.autograd.grad(y, wt, dy.float().permute(0, 3, 1, 2))
        ref, taps = gw.permute(0, 2, 3, 1).reshape(N, k * k * Cc), k * k
    dw = torch.zeros(N, taps * Cc, device=cuda)
    ops.conv_wgrad(B, H, W, x16.reshape(-1, Cc), xm, dy.reshape(-1, N), dm, dw, backend=0)
    torch.cuda.synchronize()
    assert _rel_err(dw, ref) < 2e-3


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,C,ld", [(1000, 64, 128), (777, 320, 320), (5000, 8, 24), (300, 5, 5)])
def test_colsum(cuda, M, C, ld, dt):
    """dfcsa_colsum: out[c] += sum_m x[m, c] on a channel slice of a wider buffer (bias gradients)."""
    from dfcsa import ops
    g = torch.Generator().manual_seed(29)
    buf = torch.randn(M, ld, generator=g).cuda().to(dt)
    x = buf[:, :C]
    out = torch.full((C,), 0.5, device=cuda)
    ops.colsum(x, out)
    torch.cuda.synchronize()
    assert _rel_err(out, x.float().sum(0) + 0.5) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W,C,P", [(2, 56, 56, 64, 4), (1, 40, 36, 128, 4), (2, 32, 32, 64, 8), (1, 24, 24, 256, 16),
                                       (1, 20, 28, 40, 4), (1, 7, 9, 64, 4), (1, 112, 112, 64, 4)])
def test_branch_backward_passes_match_torch(cuda, B, H, W, C, P):
    """branch_bwd_reduce1 (BN1 sums, bilinear^T of dA split over several lanes per output, d gamma) and branch_bwd_apply
    (L branch on the plain BatchNorm-backward kernel, A branch with the pool^T gather) against autograd, tensor by tensor."""
    from dfcsa import ops
    dev = cuda
    g = torch.Generator().manual_seed(11)
    M = B * H * W
    dz = torch.randn(M, 3 * C, generator=g).to(dev).bfloat16()
    l0 = torch.randn(M, C, generator=g).to(dev).half()
    a0 = torch.randn(M, C, generator=g).to(dev).half()
    o = torch.randn(B, P, P, C, generator=g).to(dev)
    dpooled = torch.randn(B, P, P, C, generator=g).to(dev)
    gamma = torch.tensor([0.7], device=dev)

    def bn(x):
        mean = x.float().mean(0)
        var = x.float().var(0, unbiased=False)
        invstd = (var + 1e-5).rsqrt()
        w = torch.rand(C, generator=g).to(dev) + 0.5
        b = 0.3 * torch.randn(C, generator=g).to(dev)
        return (w * invstd).contiguous(), (b - mean * w * invstd).contiguous(), mean.contiguous(), invstd.contiguous()

    bn1, bn2 = bn(l0), bn(a0)
    red = torch.zeros(4 * C + 2, device=dev, dtype=torch.float64)
    red1, red2, dgam = red[:2 * C], red[2 * C:4 * C], red[4 * C + 1:4 * C + 2]
    tmp = torch.empty(B, H, P, C, device=dev)
    d_o = torch.empty(B * P * P, C, device=dev)
    ops.branch_bwd_reduce1(dz, l0, None, B, H, W, bn1[0], bn1[1], bn1[2], bn1[3], None, None, o, P, gamma, red1, dgam, tmp, d_o)
    ops.branch_bwd_reduce2(dz, a0, B, H, W, bn2[0], bn2[1], bn2[2], bn2[3], dpooled, P, red2)
    dl0 = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    da0 = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    ops.branch_bwd_apply(dz, l0, a0, B, H, W, bn1, red1, bn2, red2, dpooled, P, dl0, da0)
    torch.cuda.synchronize()

    # ---- reference (fp32 autograd) ----
    dL = dz[:, C:2 * C].float()
    dA = dz[:, 2 * C:].float()
    o_n = o.permute(0, 3, 1, 2).clone().requires_grad_(True)
    U = F.interpolate(o_n, size=(H, W), mode="bilinear", align_corners=False)
    dA_n = dA.view(B, H, W, C).permute(0, 3, 1, 2)
    s = (U * dA_n).sum()
    s.backward()
    d_o_ref = (0.7 * o_n.grad).permute(0, 2, 3, 1).reshape(B * P * P, C)
    assert _rel_err(d_o, d_o_ref) < 1e-4
    assert abs(float(dgam) - float(s)) <= 1e-4 * float((U * dA_n).abs().sum()) + 1e-3

    def bn_bwd(d, x, bnp):
        sc, sh, mean, invstd = bnp
        xf = x.float()
        dm = torch.where(xf * sc + sh > 0, d, torch.zeros_like(d))
        xhat = (xf - mean) * invstd
        k1 = dm.mean(0)
        k2 = (dm * xhat).mean(0)
        return sc * (dm - k1 - xhat * k2), dm.sum(0), (dm * xhat).sum(0)

    dl0_ref, s1a, s1b = bn_bwd(dL, l0, bn1)
    assert _rel_err(red1[:C].float(), s1a) < 1e-3 and _rel_err(red1[C:].float(), s1b) < 1e-3
    assert _rel_err(dl0, dl0_ref) < 6e-3
    dp_n = dpooled.permute(0, 3, 1, 2).clone()
    x_n = torch.zeros(B, C, H, W, device=dev, requires_grad=True)
    (F.adaptive_avg_pool2d(x_n, P) * dp_n).sum().backward()
    dA_tot = dA + x_n.grad.permute(0, 2, 3, 1).reshape(M, C)
    da0_ref, s2a, s2b = bn_bwd(dA_tot, a0, bn2)
    assert _rel_err(red2[:C].float(), s2a) < 1e-3 and _rel_err(red2[C:].float(), s2b) < 1e-3
    assert _rel_err(da0, da0_ref) < 6e-3


@pytest.mark.parametrize("backend", [1, 0], ids=["simt", "tc"])
@pytest.mark.parametrize("case", [(2, 12, 12, [128], "1", 128, 64), (1, 20, 20, [64], "3", 64, 0), (2, 9, 11, [64], "1", 192, 72),
                                  (1, 16, 16, [64, 64], "31", 128, 0)])
def test_conv_gemm_bias_relu_epilogue(cuda, case, backend):
    """Inference path: bias + ReLU on the output columns n < act_cols in the GEMM epilogue (BatchNorm folded into the packed
    weights, reference inference.py:100 eval mode), both backends, against torch."""
    from dfcsa import ops
    B, H, W, cins, modes, N, act_cols = case
    g = torch.Generator().manual_seed(31)
    segs, ref, wparts = [], 0, []
    for c, m in zip(cins, modes):
        x = torch.randn(B, H, W, c, generator=g).to(cuda).half()
        k = 3 if m == "3" else 1
        wt = (torch.randn(N, c, k, k, generator=g) / (k * k * c) ** 0.5).to(cuda).half()
        ref = ref + F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), padding=k // 2)
        segs.append((x.reshape(-1, c), _MODE[m]))
        wparts.append(_pack_fwd(wt))
    w = torch.cat(wparts, dim=1).contiguous()
    if backend == ops.BACKEND_SIMT:
        w = w.float()
    bias = torch.randn(N, generator=g).to(cuda)
    ref = ref.permute(0, 2, 3, 1).reshape(-1, N) + bias
    nrelu = act_cols or N
    ref[:, :nrelu] = torch.relu(ref[:, :nrelu])
    out = torch.zeros(B * H * W, N, device=cuda, dtype=torch.float16)
    ops.conv_gemm(B, H, W, segs, w, N, out, bias=bias, backend=backend, act=1, act_cols=act_cols)
    torch.cuda.synchronize()
    assert _rel_err(out, ref) < 4e-3
    assert bool((out[:, :nrelu] >= 0).all())
    if nrelu < N:
        assert bool((out[:, nrelu:] < 0).any())           # the columns past act_cols stay linear


@pytest.mark.parametrize("B,H,W,C", [(1, 16, 16, 64), (2, 24, 40, 64), (1, 56, 56, 128), (2, 14, 14, 512), (1, 20, 28, 256)])
def test_conv_gemm_gate_mix_epilogue(cuda, B, H, W, C):
    """Inference path: the gate conv over [L | A] writes fused = s L + (1 - s) A, s = sigmoid(conv + bias), from its epilogue
    (reference models/unet_dfc_sa_res.py:104-106), into the first C columns of the same [f | L | A] buffer it reads."""
    from dfcsa import ops
    g = torch.Generator().manual_seed(41)
    M = B * H * W
    z = torch.randn(M, 3 * C, generator=g).to(cuda).half()
    z[:, :C] = 0
    wt = (torch.randn(C, 2 * C, 1, 1, generator=g) / (2 * C) ** 0.5 * 3).to(cuda).half()
    bias = torch.randn(C, generator=g).to(cuda)
    L, A = z[:, C:2 * C].float(), z[:, 2 * C:].float()
    G = z[:, C:].float() @ wt.float().reshape(C, 2 * C).t() + bias
    s = torch.sigmoid(G)
    ref = s * L + (1 - s) * A
    ops.conv_gemm(B, H, W, [(z[:, C:], 0)], _pack_fwd(wt), C, z[:, :C], bias=bias, epi=("gate_mix", z[:, C:2 * C], z[:, 2 * C:]))
    torch.cuda.synchronize()
    assert _rel_err(z[:, :C], ref) < 4e-3
    assert torch.equal(z[:, C:2 * C].float(), L) and torch.equal(z[:, 2 * C:].float(), A)      # operands untouched


@pytest.mark.parametrize("B,H,W,C", [(1, 16, 16, 64), (2, 24, 40, 64), (1, 56, 56, 128), (2, 14, 14, 512), (1, 20, 28, 1024)])
def test_conv_gemm_residual_epilogue(cuda, B, H, W, C):
    """Inference path: the fusion conv over [f | L | A] writes relu(conv + bias) + res_scale * R from its epilogue
    (reference models/unet_dfc_sa_res.py:110-114), R being the second half of the [a | R] buffer (pitch 2C) and the
    output a channel slice of a wider (concatenation) buffer."""
    from dfcsa import ops
    g = torch.Generator().manual_seed(43)
    M = B * H * W
    z = torch.randn(M, 3 * C, generator=g).to(cuda).half()
    AR = torch.randn(M, 2 * C, generator=g).to(cuda).half()
    wt = (torch.randn(C, 3 * C, 1, 1, generator=g) / (3 * C) ** 0.5).to(cuda).half()
    bias = torch.randn(C, generator=g).to(cuda)
    rs = torch.tensor([0.37], device=cuda)
    ref = torch.relu(z.float() @ wt.float().reshape(C, 3 * C).t() + bias) + 0.37 * AR[:, C:].float()
    ybuf = torch.zeros(M, 2 * C, device=cuda, dtype=torch.float16)
    ops.conv_gemm(B, H, W, [(z, 0)], _pack_fwd(wt), C, ybuf[:, C:], bias=bias, act=1, epi=("residual", AR[:, C:], rs))
    torch.cuda.synchronize()
    assert _rel_err(ybuf[:, C:], ref) < 4e-3
    assert bool((ybuf[:, :C] == 0).all())


def test_fused_epilogue_needs_the_tensor_core_backend(cuda):
    from dfcsa import ops
    z = torch.zeros(256, 192, device=cuda, dtype=torch.float16)
    w = torch.zeros(64, 128, device=cuda)
    with pytest.raises(RuntimeError, match="DFCSA_BACKEND_TC"):
        ops.conv_gemm(1, 16, 16, [(z[:, 64:], 0)], w, 64, z[:, :64], backend=ops.BACKEND_SIMT, epi=("gate_mix", z[:, 64:128], z[:, 128:]))


@pytest.mark.parametrize("backend", [1, 0], ids=["simt", "tc"])
def test_fp16_outputs_saturate_instead_of_overflowing(cuda, backend):
    """Precision policy: forward activations are stored in fp16; a result beyond +-65504 is clamped to the
    largest finite value (never inf, which would poison the next BatchNorm), and a NaN input still propagates so that the
    optimizer's non-finite check skips the step."""
    from dfcsa import ops
    B, H, W, C, N = 1, 8, 16, 64, 64
    x = torch.full((B * H * W, C), 3000.0, device=cuda, dtype=torch.float16)
    x[1::2] = -3000.0
    w = torch.ones(N, C, device=cuda, dtype=torch.float16 if backend == ops.BACKEND_TC else torch.float32)
    out = torch.zeros(B * H * W, N, device=cuda, dtype=torch.float16)
    st = torch.zeros(2 * N, device=cuda, dtype=torch.float64)
    ops.conv_gemm(B, H, W, [(x, ops.TAP_1x1)], w, N, out, stats=st, backend=backend)      # |sum| = 192000 > 65504
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out).all()) and float(out.max()) == 65504.0 and float(out.min()) == -65504.0
    assert bool(torch.isfinite(st).all())
    x[5, 3] = float("nan")
    ops.conv_gemm(B, H, W, [(x, ops.TAP_1x1)], w, N, out, backend=backend)
    torch.cuda.synchronize()
    assert bool(torch.isnan(out[5]).all()) and bool(torch.isfinite(out[6]).all())
