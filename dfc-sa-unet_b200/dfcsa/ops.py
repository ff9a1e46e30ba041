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
              backend=BACKEND_TC):
    """out[m, n] (+)= sum_seg sum_tap sum_c seg[pix(m,tap), c] * w[n, k]   (dfcsa_conv_gemm)."""
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
    L.check(L.lib().dfcsa_conv_gemm(C.byref(p), backend, L.stream()), "dfcsa_conv_gemm")


def conv_wgrad(B, H, W, x, x_mode, dy, dy_mode, dw, alpha=None, backend=BACKEND_TC):
    """dw[n, t*C + c] += alpha * sum_m dy[., n] * x[., c]   (dfcsa_conv_wgrad); dw is a 2-D fp32 view."""
    p = L.WgradParams()
    p.B, p.H, p.W = B, H, W
    p.x, p.ld_x, p.C, p.x_dtype, p.x_tap_mode = x.data_ptr(), _mat(x), x.shape[1], L.dt(x), x_mode
    p.dy, p.ld_dy, p.N, p.dy_dtype, p.dy_tap_mode = dy.data_ptr(), _mat(dy), dy.shape[1], L.dt(dy), dy_mode
    p.dw, p.ld_dw = dw.data_ptr(), dw.stride(0)
    p.alpha = alpha.data_ptr() if alpha is not None else None
    L.check(L.lib().dfcsa_conv_wgrad(C.byref(p), backend, L.stream()), "dfcsa_conv_wgrad")


def wgrad_tc_eligible(x, dy):
    return (x.dtype != torch.float32 and dy.dtype != torch.float32 and x.shape[1] % 64 == 0 and dy.shape[1] % 8 == 0
            and x.stride(0) % 8 == 0 and dy.stride(0) % 8 == 0 and x.data_ptr() % 16 == 0 and dy.data_ptr() % 16 == 0)


def permute3(src, dst, D, s, flip1=False, scale=None):
    """dst[(i0*D1+i1)*D2+i2] = scale*src[i0*s0 + i1'*s1 + i2*s2]   (dfcsa_permute3)."""
    L.check(L.lib().dfcsa_permute3(L.ptr(src), L.dt(src), L.ptr(dst), L.dt(dst),
                                   C.c_int64(D[0]), C.c_int64(D[1]), C.c_int64(D[2]),
                                   C.c_int64(s[0]), C.c_int64(s[1]), C.c_int64(s[2]),
                                   1 if flip1 else 0, L.ptr(scale), L.stream()), "dfcsa_permute3")


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
    L.check(L.lib().dfcsa_sgemm(C.byref(p), L.stream()), "dfcsa_sgemm")


def softmax_rows(x, y):
    rows, cols = x.numel() // x.shape[-1], x.shape[-1]
    L.check(L.lib().dfcsa_softmax_rows(L.ptr(x), L.ptr(y), C.c_int64(rows), cols, L.stream()), "dfcsa_softmax_rows")


def softmax_rows_bwd(y, dy, dx):
    rows, cols = y.numel() // y.shape[-1], y.shape[-1]
    L.check(L.lib().dfcsa_softmax_rows_bwd(L.ptr(y), L.ptr(dy), L.ptr(dx), C.c_int64(rows), cols, L.stream()),
            "dfcsa_softmax_rows_bwd")
