"""Thin Python wrappers over the C ABI (one function per entry point of include/dfcsa.h).

Activations are 2-D "pixel-major" views [M, C] with stride (ld, 1) of NHWC storage; a channel slice of a wider
buffer is just `buf[:, a:b]`.  Nothing here computes: every function enqueues one libdfcsa kernel on the current
CUDA stream.
"""
import ctypes as C

import torch

from . import _lib as L
from ._lib import (BACKEND_SIMT, BACKEND_TC, BF16, F16, F32, OUT_CONVT2x2, OUT_DIRECT, TAP_1x1, TAP_2x2S2, TAP_3x3)


def _mat(t):
    assert t.dim() == 2 and (t.stride(1) == 1 or t.shape[1] == 1), "expected a [M, C] view with unit channel stride"
    return t.stride(0)


def tc_eligible(segs, N, out):
    """The tcgen05 path takes 16-bit operands whose channel counts are multiples of 64."""
    for m, _ in segs:
        if m.dtype == torch.float32 or m.shape[1] % 64 != 0 or m.stride(0) % 8 != 0 or m.data_ptr() % 16 != 0:
            return False
    return N % 8 == 0 and out.stride(0) % 8 == 0 and out.data_ptr() % 16 == 0


def conv_gemm(B, H, W, segs, w, N, out, out_mode=OUT_DIRECT, accumulate=False, bias=None, stats=None,
              backend=BACKEND_TC, shadow=None, act=0, act_cols=0, stats_cols=0, bn=None, epi=None):
    """out[m, n] (+)= sum_seg sum_tap sum_c seg[pix(m,tap), c] * w[n, k]   (dfcsa_conv_gemm).
    epi = ("gate_mix", p, q): out = s p + (1 - s) q with s = sigmoid(result);  ("residual", p, scale): out = result + scale p
    (fused inference epilogues, tensor-core backend only)."""
    p = L.ConvParams()
    p.B, p.H, p.W, p.n_seg = B, H, W, len(segs)
    for i, (m, mode) in enumerate(segs):
        p.seg[i].ptr = m.data_ptr()
        p.seg[i].ld = _mat(m)
        p.seg[i].channels = m.shape[1]
        p.seg[i].tap_mode = mode
    p.src_dtype = L.dt(segs[0][0])
    p.w_dtype = L.dt(w)
    p.w = w.data_ptr()
    p.N = N
    p.out_dtype = L.dt(out)
    p.out = out.data_ptr()
    p.ld_out = _mat(out)
    p.out_mode = out_mode
    p.accumulate = 1 if accumulate else 0
    p.bias = bias.data_ptr() if bias is not None else None
    p.stats = stats.data_ptr() if stats is not None else None
    p.shadow = shadow.data_ptr() if shadow is not None else None
    p.ld_shadow = _mat(shadow) if shadow is not None else 0
    p.act, p.act_cols, p.stats_cols = act, act_cols, stats_cols
    if bn is not None:      # L.BnFold: BatchNorm finalize inside this launch (kept alive until the call returns)
        p.bn = C.addressof(bn)
    if epi is not None:
        e = L.ConvEpi()
        if epi[0] == "gate_mix":
            e.mode, e.p, e.q, e.ld = L.EPI_GATE_MIX, epi[1].data_ptr(), epi[2].data_ptr(), _mat(epi[1])
            assert _mat(epi[2]) == e.ld and epi[1].dtype == epi[2].dtype == torch.float16
        else:
            e.mode, e.p, e.ld, e.scale = L.EPI_RESIDUAL, epi[1].data_ptr(), _mat(epi[1]), epi[2].data_ptr()
            assert epi[1].dtype == torch.float16 and epi[2].dtype == torch.float32
        p.epi = C.addressof(e)
    ktot = sum(m.shape[1] * (1 if mode == TAP_1x1 else 9 if mode == TAP_3x3 else 4) for m, mode in segs)
    L.call("dfcsa_conv_gemm", C.byref(p), backend, L.stream(), tag="conv_tc" if backend == BACKEND_TC else "conv_simt",
           flops=2.0 * B * H * W * N * ktot,
           desc=f"M={B * H * W} {H}x{W} N={N} K={ktot} segs={[(m.shape[1], mode) for m, mode in segs]} acc={int(bool(accumulate))} "
                f"stats={int(stats is not None)} out={str(out.dtype)[6:]} mode={out_mode}")


def conv_wgrad(B, H, W, x, x_mode, dy, dy_mode, dw, alpha=None, backend=BACKEND_TC, second=None):
    """dw[n, t*C + c] += alpha * sum_m dy[., n] * x[., c]   (dfcsa_conv_wgrad); dw is a 2-D fp32 view.
    second = (dy2, dw2, c_begin2, alpha2): a second 1x1 gradient over the same x in the same launch."""
    p = L.WgradParams()
    p.B, p.H, p.W = B, H, W
    p.x, p.ld_x, p.C, p.x_dtype, p.x_tap_mode = x.data_ptr(), _mat(x), x.shape[1], L.dt(x), x_mode
    p.dy, p.ld_dy, p.N, p.dy_dtype, p.dy_tap_mode = dy.data_ptr(), _mat(dy), dy.shape[1], L.dt(dy), dy_mode
    p.dw, p.ld_dw = dw.data_ptr(), dw.stride(0)
    p.alpha = alpha.data_ptr() if alpha is not None else None
    flops2 = 0.0
    if second is not None:
        dy2, dw2, c_begin2, alpha2 = second
        p.dy2, p.ld_dy2, p.N2, p.c_begin2 = dy2.data_ptr(), _mat(dy2), dy2.shape[1], c_begin2
        p.dw2, p.ld_dw2 = dw2.data_ptr(), dw2.stride(0)
        p.alpha2 = alpha2.data_ptr() if alpha2 is not None else None
        flops2 = 2.0 * B * H * W * (x.shape[1] - c_begin2) * dy2.shape[1]
    taps = 9 if x_mode == TAP_3x3 else (4 if dy_mode == TAP_2x2S2 else 1)
    L.call("dfcsa_conv_wgrad", C.byref(p), backend, L.stream(), tag="wgrad_tc" if backend == BACKEND_TC else "wgrad_simt",
           flops=2.0 * B * H * W * taps * x.shape[1] * dy.shape[1] + flops2,
           desc=f"M={B * H * W} {H}x{W} N={dy.shape[1]}{'+' + str(second[0].shape[1]) if second is not None else ''} C={x.shape[1]} taps={taps}")


def wgrad_tc_eligible(x, dy):
    return (x.dtype != torch.float32 and dy.dtype != torch.float32 and x.shape[1] % 64 == 0 and dy.shape[1] % 8 == 0
            and x.stride(0) % 8 == 0 and dy.stride(0) % 8 == 0 and x.data_ptr() % 16 == 0 and dy.data_ptr() % 16 == 0)


def permute3(src, dst, D, s, flip1=False, scale=None, ld_dst=0):
    """dst[i0*ld_dst + i1*D2 + i2] = scale*src[i0*s0 + i1'*s1 + i2*s2]   (dfcsa_permute3); ld_dst=0: dense."""
    L.call("dfcsa_permute3", L.ptr(src), L.dt(src), L.ptr(dst), L.dt(dst),
                                   C.c_int64(D[0]), C.c_int64(D[1]), C.c_int64(D[2]),
                                   C.c_int64(s[0]), C.c_int64(s[1]), C.c_int64(s[2]),
                                   1 if flip1 else 0, L.ptr(scale), C.c_int64(ld_dst), L.stream())


def pack_jobs(table, n_jobs, prefix, total_chunks):
    """run a PackPlan's device job table (dfcsa_pack_jobs): every weight re-layout of the network in one launch."""
    L.call("dfcsa_pack_jobs", L.ptr(table), n_jobs, L.ptr(prefix), C.c_int64(total_chunks), L.stream())


def sgemm(batch, M, N, K, A, a_str, Bm, b_str, Cm, c_str, alpha=1.0, beta=0.0, bias_n=None, bias_m=None):
    """C[b] = alpha*A[b]@B[b] + beta*C[b] with element strides (batch, row, col)   (dfcsa_sgemm)."""
    p = L.SgemmParams()
    p.batch, p.M, p.N, p.K = batch, M, N, K
    p.A, (p.a_b, p.a_m, p.a_k) = A.data_ptr(), a_str
    p.B, (p.b_b, p.b_k, p.b_n) = Bm.data_ptr(), b_str
    p.C, (p.c_b, p.c_m, p.c_n) = Cm.data_ptr(), c_str
    p.bias_n = bias_n.data_ptr() if bias_n is not None else None
    p.bias_m = bias_m.data_ptr() if bias_m is not None else None
    p.alpha, p.beta = alpha, beta
    L.call("dfcsa_sgemm", C.byref(p), L.stream())


def softmax_rows(x, y):
    """y = softmax(x) along the last dim; x fp32, y fp32 / fp16 / bf16."""
    rows, cols = x.numel() // x.shape[-1], x.shape[-1]
    L.call("dfcsa_softmax_rows", L.ptr(x), L.ptr(y), L.dt(y), C.c_int64(rows), cols, L.stream())


def softmax_rows_bwd(y, dy, dx):
    """dx = y * (dy - rowsum(dy * y)); y fp32 / 16-bit, dy fp32, dx fp32 / 16-bit."""
    rows, cols = y.numel() // y.shape[-1], y.shape[-1]
    L.call("dfcsa_softmax_rows_bwd", L.ptr(y), L.dt(y), L.ptr(dy), L.ptr(dx), L.dt(dx), C.c_int64(rows), cols, L.stream())


def softmax_rows_bwd_d(y, dy, D, dx):
    """dx = y * (dy - D[row]) in one pass (dfcsa_softmax_rows_bwd_d); D from rowdot(d_o, o)."""
    rows, cols = y.numel() // y.shape[-1], y.shape[-1]
    L.call("dfcsa_softmax_rows_bwd_d", L.ptr(y), L.dt(y), L.ptr(dy), L.dt(dy), L.ptr(D), L.ptr(dx), L.dt(dx), C.c_int64(rows), cols,
           L.stream())


def rowdot(a, b, out):
    rows, cols = a.numel() // a.shape[-1], a.shape[-1]
    L.call("dfcsa_rowdot", L.ptr(a), L.ptr(b), C.c_int64(rows), cols, L.ptr(out), L.stream())


def attn_small_fwd(qkv, B, N, Cq, Cn, attn, o):
    L.call("dfcsa_attn_small_fwd", L.ptr(qkv), _i64(qkv.stride(0)), B, N, Cq, Cn, L.ptr(attn), L.ptr(o), L.stream())


def attn_small_bwd(qkv, attn, d_o, B, N, Cq, Cn, dqkv, dbq=None, dbk=None, dbv=None):
    """dbq / dbk / dbv: optional bias-gradient tensors the kernel ACCUMULATES the column sums of dq / dk / dv into."""
    L.call("dfcsa_attn_small_bwd", L.ptr(qkv), _i64(qkv.stride(0)), L.ptr(attn), L.ptr(d_o), B, N, Cq, Cn, L.ptr(dqkv),
           L.ptr(dbq), L.ptr(dbk), L.ptr(dbv), L.stream())


def softmax_bgemm(batch, M, N, K, A, a_b, ld_a, Bm, b_b, ld_b, out, lse=None, have_lse=False):
    """out[b] = softmax_rows(A[b] @ B[b]^T) (both operands K-major) in two tcgen05 launches that never write the logits:
    a row-statistics pass (ROWSTATS epilogue + dfcsa_lse_combine -> lse [batch*M] fp32) and a pass whose epilogue stores
    exp(x - lse).  With have_lse the statistics pass is skipped (the backward recompute reuses the forward's lse)."""
    dev = A.device
    if lse is None:
        assert not have_lse
        lse = torch.empty((batch * M,), dtype=torch.float32, device=dev)
    if not have_lse:
        parts = L.lib().dfcsa_bgemm_rowstat_parts(N)
        rowstat = torch.empty((batch, parts, M, 2), dtype=torch.float32, device=dev)
        bgemm(batch, M, N, K, A, a_b, ld_a, False, Bm, b_b, ld_b, False, None, 0, 0, epi=1, rowstat=rowstat)
        L.call("dfcsa_lse_combine", L.ptr(rowstat), parts, batch, M, L.ptr(lse), L.stream())
    bgemm(batch, M, N, K, A, a_b, ld_a, False, Bm, b_b, ld_b, False, out, M * N, N, epi=2, rowvec=lse)


def attn_row_lse(batch, N, Cq, qkv, ld, lse):
    """lse[b*N + i] = log sum_j exp(q_i . k_j) for fp16 rows (q | k | v) of pitch ld: the statistics pass alone."""
    parts = L.lib().dfcsa_bgemm_rowstat_parts(N)
    rowstat = torch.empty((batch, parts, N, 2), dtype=torch.float32, device=qkv.device)
    bgemm(batch, N, N, Cq, qkv[:, :Cq], N * ld, ld, False, qkv[:, Cq:2 * Cq], N * ld, ld, False, None, 0, 0, epi=1, rowstat=rowstat)
    L.call("dfcsa_lse_combine", L.ptr(rowstat), parts, batch, N, L.ptr(lse), L.stream())


def attn_pv_fused(qkv, batch, N, Cq, Cn, lse, o):
    """o[b] = exp(q k^T - lse) v in ONE tcgen05 kernel, the probabilities never leave the SM (dfcsa_attn_pv_fused)."""
    assert qkv.dtype == torch.float16 and qkv.stride(1) == 1 and o.is_contiguous() and o.dtype == torch.float32
    L.call("dfcsa_attn_pv_fused", L.ptr(qkv), _i64(qkv.stride(0)), batch, N, Cq, Cn, L.ptr(lse), L.ptr(o), L.stream(),
           tag="attn_pv_fused", flops=2.0 * batch * N * N * (Cq + Cn), desc=f"b={batch} N={N} Cq={Cq} C={Cn}")


def attn_bwd_fused(qkv16, qkvb, dO, batch, N, Cq, Cn, lse, D, dqkv):
    """dq | dk | dv (fp32 rows of dqkv) from the fp16 forward operands, their bf16 copies, dO (bf16), the forward's row
    log-sum-exp and D = rowdot(dO, O): two tcgen05 kernels, no [N, N] tensor in HBM (dfcsa_attn_bwd_fused)."""
    assert qkv16.dtype == torch.float16 and qkvb.dtype == torch.bfloat16 and dO.dtype == torch.bfloat16 and dqkv.dtype == torch.float32
    L.call("dfcsa_attn_bwd_fused", L.ptr(qkv16), _i64(qkv16.stride(0)), L.ptr(qkvb), _i64(qkvb.stride(0)), L.ptr(dO), _i64(dO.stride(0)),
           batch, N, Cq, Cn, L.ptr(lse), L.ptr(D), L.ptr(dqkv), _i64(dqkv.stride(0)), L.stream(),
           tag="attn_bwd_fused", flops=2.0 * batch * N * N * (3 * Cq + 2 * Cn + Cq + Cn), desc=f"b={batch} N={N} Cq={Cq} C={Cn}")


def softmax_bwd_bgemm(batch, M, N, K, dO, a_b, ld_a, V, b_b, ld_b, probs, D, dS):
    """dS[b] = probs[b] * (dO[b] @ V[b]^T - D[b][:, None]) (both operands K-major): the softmax backward fused into the
    epilogue of the dP product, D = rowdot(dO, O).  dS may be probs itself (same element size)."""
    assert probs.shape == dS.shape and probs.is_contiguous() and dS.is_contiguous() and D.dtype == torch.float32
    bgemm(batch, M, N, K, dO, a_b, ld_a, False, V, b_b, ld_b, False, dS, M * N, N, epi=3, rowvec=D, aux=probs)


def bgemm(batch, M, N, K, A, a_b, ld_a, a_mn, Bm, b_b, ld_b, b_mn, Cm, c_b, ld_c, epi=0, rowstat=None, rowvec=None, aux=None):
    """C[b] = A[b] @ B[b] on tcgen05 (dfcsa_bgemm); operands 16-bit of one dtype, K-major or MN-major (see dfcsa.h)."""
    p = L.BgemmParams()
    p.epi_mode = epi
    p.rowstat = rowstat.data_ptr() if rowstat is not None else None
    p.rowvec = rowvec.data_ptr() if rowvec is not None else None
    if aux is not None:
        p.aux, p.aux_dtype = aux.data_ptr(), L.dt(aux)
    p.batch, p.M, p.N, p.K = batch, M, N, K
    p.A, p.a_b, p.ld_a, p.a_mn_major = A.data_ptr(), a_b, ld_a, 1 if a_mn else 0
    p.B, p.b_b, p.ld_b, p.b_mn_major = Bm.data_ptr(), b_b, ld_b, 1 if b_mn else 0
    p.ab_dtype = L.dt(A)
    assert A.dtype == Bm.dtype
    if Cm is not None:
        p.C, p.c_b, p.ld_c, p.c_dtype = Cm.data_ptr(), c_b, ld_c, L.dt(Cm)
    L.call("dfcsa_bgemm", C.byref(p), L.stream(), tag="bgemm_tc", flops=2.0 * batch * M * N * K,
           desc=f"b={batch} M={M} N={N} K={K} a_mn={int(a_mn)} b_mn={int(b_mn)}")


def _i64(v):
    return C.c_int64(int(v))


def bn_finalize(s_sum, s_sq, count, gamma, beta, conv_bias, rmean, rvar, momentum, eps, scale, shift, mean, invstd):
    Cn = gamma.numel()
    L.call("dfcsa_bn_finalize", L.ptr(s_sum), L.ptr(s_sq), _i64(count), Cn, L.ptr(gamma), L.ptr(beta),
                                      L.ptr(conv_bias), L.ptr(rmean), L.ptr(rvar), C.c_float(momentum), C.c_float(eps),
                                      L.ptr(scale), L.ptr(shift), L.ptr(mean), L.ptr(invstd), L.stream())


def bn_fold(gamma, beta, conv_bias, rmean, rvar, momentum, eps, count, aff, ticket):
    """dfcsa_bn_fold_t for conv_gemm(bn=...): aff is the [4, C] (scale, shift, mean, invstd) tensor, ticket a zeroed int32 [1]."""
    b = L.BnFold()
    b.gamma, b.beta = gamma.data_ptr(), beta.data_ptr()
    b.conv_bias = conv_bias.data_ptr() if conv_bias is not None else None
    b.running_mean, b.running_var = rmean.data_ptr(), rvar.data_ptr()
    b.momentum, b.eps, b.count = momentum, eps, count
    b.scale, b.shift, b.mean, b.invstd = (aff[i].data_ptr() for i in range(4))
    b.ticket, b.channels = ticket.data_ptr(), gamma.numel()
    return b


def bn_eval_affine(gamma, beta, conv_bias, rmean, rvar, eps, scale, shift):
    L.call("dfcsa_bn_eval_affine", gamma.numel(), L.ptr(gamma), L.ptr(beta), L.ptr(conv_bias), L.ptr(rmean),
                                         L.ptr(rvar), C.c_float(eps), L.ptr(scale), L.ptr(shift), L.stream())


def bn_param_grads(red, Cn, dgamma, dbeta):
    L.call("dfcsa_bn_param_grads", L.ptr(red), Cn, L.ptr(dgamma), L.ptr(dbeta), L.stream())


def block_param_grads(red, Cn, bn_grads, drs=None, dgam=None):
    """all small parameter gradients of one block from its reduction buffer in one launch (dfcsa_block_param_grads).
    bn_grads: four (dgamma, dbeta) pairs or None, in the order of the buffer's four BatchNorm slots."""
    args = []
    for pair in bn_grads:
        args += [L.ptr(pair[0]), L.ptr(pair[1])] if pair is not None else [None, None]
    L.call("dfcsa_block_param_grads", L.ptr(red), Cn, *args, L.ptr(drs), L.ptr(dgam), L.stream())


def bnrelu_pool_fwd(a0, B, H, W, scale, shift, P, tmp, pooled, with_masks=False):
    """with_masks: tmp [3, B, H, P, C] and pooled [3, B*P*P, C] - plane 0 the pooled activation, planes 1 / 2 the window means
    of the ReLU mask and of mask * A0 (for pool_window_terms in the backward pass)."""
    L.call("dfcsa_bnrelu_pool_fwd", L.ptr(a0), _i64(_mat(a0)), B, H, W, a0.shape[1], L.ptr(scale), L.ptr(shift), P,
                                          L.ptr(tmp), L.ptr(pooled), 1 if with_masks else 0, L.stream())


def bn_bwd_reduce(dy, x, scale, shift, mean, invstd, red):
    M, Cn = x.shape
    L.call("dfcsa_bn_bwd_reduce", L.ptr(dy), _i64(_mat(dy)), L.ptr(x), _i64(_mat(x)), _i64(M), Cn, L.ptr(scale), L.ptr(shift),
           L.ptr(mean), L.ptr(invstd), L.ptr(red), L.stream())


def pool_window_terms(dpooled, means, B, P, Cn, mean, invstd, red):
    L.call("dfcsa_pool_window_terms", L.ptr(dpooled), L.ptr(means), B, P, Cn, L.ptr(mean), L.ptr(invstd), L.ptr(red), L.stream())


def branch_act_fwd(l0, a0, B, H, W, s1, t1, s2, t2, o, P, gamma, z, zb=None):
    Cn = a0.shape[1]
    L.call("dfcsa_branch_act_fwd", L.ptr(l0), _i64(_mat(l0) if l0 is not None else 0), L.ptr(a0), _i64(_mat(a0)), B, H, W, Cn, L.ptr(s1),
                                         L.ptr(t1), L.ptr(s2), L.ptr(t2), L.ptr(o), P, L.ptr(gamma), L.ptr(z),
                                         _i64(_mat(z)), L.ptr(zb), _i64(_mat(zb) if zb is not None else 0), L.stream())


def gate_mix_fwd(g0, s3, t3, z, zb=None):
    M, Cn = g0.shape
    L.call("dfcsa_gate_mix_fwd", L.ptr(g0), _i64(_mat(g0)), _i64(M), Cn, L.ptr(s3), L.ptr(t3), L.ptr(z),
                                       _i64(_mat(z)), L.ptr(zb), _i64(_mat(zb) if zb is not None else 0), L.stream())


def block_out_fwd(f0, r, B, H, W, s4, t4, res_scale, y, yp=None, yb=None, ypb=None):
    Cn = f0.shape[1]
    L.call("dfcsa_block_out_fwd", L.ptr(f0), _i64(_mat(f0)), L.ptr(r), _i64(_mat(r)), B, H, W, Cn, L.ptr(s4),
                                        L.ptr(t4), L.ptr(res_scale), L.ptr(y), _i64(_mat(y)), L.ptr(yp),
                                        _i64(_mat(yp) if yp is not None else 0), L.ptr(yb), _i64(_mat(yb) if yb is not None else 0),
                                        L.ptr(ypb), _i64(_mat(ypb) if ypb is not None else 0), L.stream())


def sum_out_fwd(a, b, r, B, H, W, res_scale, y, yp=None):
    """y = a [+ b] + res_scale * r (+ 2x2 max pool into yp): the plain-sum output stage of the ablation blocks."""
    Cn = a.shape[1]
    L.call("dfcsa_sum_out_fwd", L.ptr(a), _i64(_mat(a)), L.ptr(b), _i64(_mat(b) if b is not None else 0), L.ptr(r), _i64(_mat(r)),
           B, H, W, Cn, L.ptr(res_scale), L.ptr(y), _i64(_mat(y)), L.ptr(yp), _i64(_mat(yp) if yp is not None else 0), L.stream())


def block_out_bwd_reduce(dskip, dyp, y, f0, r, B, H, W, s4, t4, mean4, invstd4, dy_out, red4, drs):
    Cn = f0.shape[1]
    L.call("dfcsa_block_out_bwd_reduce", 
        L.ptr(dskip), _i64(_mat(dskip) if dskip is not None else 0), L.ptr(dyp), _i64(_mat(dyp) if dyp is not None else 0),
        L.ptr(y), _i64(_mat(y)), L.ptr(f0), _i64(_mat(f0)), L.ptr(r), _i64(_mat(r)), B, H, W, Cn,
        L.ptr(s4), L.ptr(t4), L.ptr(mean4), L.ptr(invstd4), L.ptr(dy_out), _i64(_mat(dy_out) if dy_out is not None else 0),
        L.ptr(red4), L.ptr(drs), L.stream())


def bn_bwd_apply(dy, x, scale, shift, mean, invstd, red, act_mode, dx):
    M, Cn = x.shape
    L.call("dfcsa_bn_bwd_apply", L.ptr(dy), _i64(_mat(dy)), L.ptr(x), _i64(_mat(x)), _i64(M), Cn, L.ptr(scale),
                                       L.ptr(shift), L.ptr(mean), L.ptr(invstd), None, L.ptr(red), act_mode, L.ptr(dx),
                                       _i64(_mat(dx)), L.stream())


def gate_mix_bwd_reduce(dz, z, g0, s3, t3, mean3, invstd3, red3):
    M, Cn = g0.shape
    L.call("dfcsa_gate_mix_bwd_reduce", L.ptr(dz), _i64(_mat(dz)), L.ptr(z), _i64(_mat(z)), L.ptr(g0), _i64(_mat(g0)),
                                              _i64(M), Cn, L.ptr(s3), L.ptr(t3), L.ptr(mean3), L.ptr(invstd3), L.ptr(red3),
                                              L.stream())


def gate_mix_bwd_apply(dz, z, g0, s3, t3, mean3, invstd3, red3, dg0):
    M, Cn = g0.shape
    L.call("dfcsa_gate_mix_bwd_apply", L.ptr(dz), _i64(_mat(dz)), L.ptr(z), _i64(_mat(z)), L.ptr(g0), _i64(_mat(g0)),
                                             _i64(M), Cn, L.ptr(s3), L.ptr(t3), L.ptr(mean3), L.ptr(invstd3), None,
                                             L.ptr(red3), L.ptr(dg0), _i64(_mat(dg0)), L.stream())


def branch_bwd_reduce1(dz, l0, g0, B, H, W, s1, t1, mean1, invstd1, s3, t3, o, P, gamma, red1, dgamma, tmp, d_o):
    Cn = l0.shape[1]
    L.call("dfcsa_branch_bwd_reduce1", L.ptr(dz), _i64(_mat(dz)), L.ptr(l0), _i64(_mat(l0)), L.ptr(g0), _i64(_mat(g0) if g0 is not None else 0),
                                             B, H, W, Cn, L.ptr(s1),
                                             L.ptr(t1), L.ptr(mean1), L.ptr(invstd1), L.ptr(s3), L.ptr(t3), L.ptr(o), P, L.ptr(gamma), L.ptr(red1),
                                             L.ptr(dgamma), L.ptr(tmp), L.ptr(d_o), L.stream())


def branch_bwd_reduce2(dz, a0, B, H, W, s2, t2, mean2, invstd2, dpooled, P, red2):
    Cn = a0.shape[1]
    L.call("dfcsa_branch_bwd_reduce2", L.ptr(dz), _i64(_mat(dz)), L.ptr(a0), _i64(_mat(a0)), B, H, W, Cn, L.ptr(s2),
                                             L.ptr(t2), L.ptr(mean2), L.ptr(invstd2), L.ptr(dpooled), P, L.ptr(red2),
                                             L.stream())


def branch_bwd_apply(dz, l0, a0, B, H, W, bn1, red1, bn2, red2, dpooled, P, dl0, da0):
    Cn = l0.shape[1]
    s1, t1, m1, i1 = bn1
    s2, t2, m2, i2 = bn2
    L.call("dfcsa_branch_bwd_apply", 
        L.ptr(dz), _i64(_mat(dz)), L.ptr(l0), _i64(_mat(l0)), L.ptr(a0), _i64(_mat(a0)), B, H, W, Cn,
        L.ptr(s1), L.ptr(t1), L.ptr(m1), L.ptr(i1), None, L.ptr(red1),
        L.ptr(s2), L.ptr(t2), L.ptr(m2), L.ptr(i2), None, L.ptr(red2),
        L.ptr(dpooled), P, L.ptr(dl0), _i64(_mat(dl0)), L.ptr(da0), _i64(_mat(da0)), L.stream())


def nchw_to_nhwc(src, dst2d, B, Cn, H, W):
    L.call("dfcsa_nchw_to_nhwc", L.ptr(src), L.ptr(dst2d), L.dt(dst2d), _i64(_mat(dst2d)), B, Cn, H, W, L.stream())


def nhwc_to_nchw(src2d, dst, B, Cn, H, W):
    L.call("dfcsa_nhwc_to_nchw", L.ptr(src2d), L.dt(src2d), _i64(_mat(src2d)), L.ptr(dst), B, Cn, H, W, L.stream())


def colsum(x2d, out):
    M, Cn = x2d.shape
    L.call("dfcsa_colsum", L.ptr(x2d), L.dt(x2d), _i64(_mat(x2d)), _i64(M), Cn, L.ptr(out), L.stream())


def cast2d(x2d, y2d):
    M, Cn = x2d.shape
    L.call("dfcsa_cast2d", L.ptr(x2d), L.dt(x2d), _i64(_mat(x2d)), L.ptr(y2d), L.dt(y2d), _i64(_mat(y2d)), _i64(M), Cn,
                                 L.stream())


def bce_dice_sums(x, t, from_logits, sums):
    L.call("dfcsa_bce_dice_sums", L.ptr(x), L.ptr(t), _i64(x.numel()), 1 if from_logits else 0, L.ptr(sums), L.stream())


def bce_dice_finalize(sums, n, w_bce, w_dice, smooth, out):
    L.call("dfcsa_bce_dice_finalize", L.ptr(sums), _i64(n), C.c_float(w_bce), C.c_float(w_dice), C.c_float(smooth),
                                            L.ptr(out), L.stream())


def bce_dice_per_sample(x, t, from_logits, w_bce, w_dice, smooth, sums, out):
    """x, t: [n, ...] contiguous fp32; sums: zeroed double [n, 8]; out: fp32 [n, 5] = (loss, bce, dice loss, hard IoU, hard Dice)
    of every sample on its own (dfcsa_bce_dice_sums_batched + dfcsa_bce_dice_finalize_batched)."""
    n = x.shape[0]
    per = x.numel() // n
    L.call("dfcsa_bce_dice_sums_batched", L.ptr(x), L.ptr(t), _i64(per), n, 1 if from_logits else 0, L.ptr(sums), L.stream())
    L.call("dfcsa_bce_dice_finalize_batched", L.ptr(sums), _i64(per), n, C.c_float(w_bce), C.c_float(w_dice), C.c_float(smooth),
           L.ptr(out), L.stream())


def bce_dice_bwd(x, t, from_logits, sums, w_bce, w_dice, smooth, gout, dx):
    L.call("dfcsa_bce_dice_bwd", L.ptr(x), L.ptr(t), _i64(x.numel()), 1 if from_logits else 0, L.ptr(sums),
                                       C.c_float(w_bce), C.c_float(w_dice), C.c_float(smooth), L.ptr(gout), L.ptr(dx),
                                       L.dt(dx), L.stream())


def grad_sumsq(table, n_tensors, max_n, sumsq):
    L.call("dfcsa_grad_sumsq", L.ptr(table), n_tensors, _i64(max_n), L.ptr(sumsq), L.stream())


def sgd_step(table, n_tensors, max_n, sumsq, gscale, max_norm, lr, momentum, weight_decay, first_step):
    L.call("dfcsa_sgd_step", L.ptr(table), n_tensors, _i64(max_n), L.ptr(sumsq), C.c_float(gscale), C.c_float(max_norm),
                                   C.c_float(lr), C.c_float(momentum), C.c_float(weight_decay), 1 if first_step else 0,
                                   L.stream())


def resize_bilinear(src, B, Hi, Wi, dst, Ho, Wo, alpha=None, add=None):
    """dst = alpha * bilinear(src: Hi x Wi -> Ho x Wo) + add on [M, C] NHWC views (dfcsa_resize_bilinear)."""
    L.call("dfcsa_resize_bilinear", L.ptr(src), L.dt(src), _i64(_mat(src)), B, Hi, Wi, src.shape[1], L.ptr(dst), L.dt(dst), _i64(_mat(dst)),
           Ho, Wo, L.ptr(alpha), L.ptr(add), L.dt(add) if add is not None else 0, _i64(_mat(add) if add is not None else 0), L.stream())


def resize_bilinear_bwd(ddst, B, Hi, Wi, dsrc, Ho, Wo, alpha=None):
    """dsrc = alpha * bilinear^T(ddst)  (dfcsa_resize_bilinear_bwd); ddst on the Ho x Wo grid, dsrc on Hi x Wi."""
    L.call("dfcsa_resize_bilinear_bwd", L.ptr(ddst), L.dt(ddst), _i64(_mat(ddst)), B, Hi, Wi, dsrc.shape[1], L.ptr(dsrc), L.dt(dsrc),
           _i64(_mat(dsrc)), Ho, Wo, L.ptr(alpha), L.stream())


def adaptive_pool(src, B, H, W, P, pooled):
    """pooled [B*P*P, C] fp32 = adaptive_avg_pool2d(src, P)  (dfcsa_adaptive_pool)."""
    L.call("dfcsa_adaptive_pool", L.ptr(src), L.dt(src), _i64(_mat(src)), B, H, W, src.shape[1], P, L.ptr(pooled), L.stream())


def adaptive_pool_bwd(dpooled, B, H, W, P, dst, add=None):
    """dst = add + adaptive_avg_pool2d^T(dpooled)  (dfcsa_adaptive_pool_bwd)."""
    L.call("dfcsa_adaptive_pool_bwd", L.ptr(dpooled), B, H, W, dst.shape[1], P, L.ptr(add), L.dt(add) if add is not None else 0,
           _i64(_mat(add) if add is not None else 0), L.ptr(dst), L.dt(dst), _i64(_mat(dst)), L.stream())


def accumulate(dst, src):
    """dst += src on flat fp32 buffers (dfcsa_accumulate)."""
    assert dst.dtype == torch.float32 and src.dtype == torch.float32 and dst.numel() == src.numel()
    L.call("dfcsa_accumulate", L.ptr(dst), L.ptr(src), _i64(dst.numel()), L.stream())
