#!/usr/bin/env python
"""The five batched-GEMM launches of one full-resolution attention block (N = H*W keys, one image), timed with CUDA
events: row statistics, exp, P V (separately and as the fused forward kernel), P^T dO, softmax-backward dS, dS K, dS^T Q.  Run under ncu for the per-kernel picture.

  python tools/attn_probe.py [--n 50176] [--c 64] [--cq 8] [--reps 3]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dfc-sa-unet_b200")]
import torch  # noqa: E402
from dfcsa import _lib as L, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=50176)
    ap.add_argument("--c", type=int, default=64)
    ap.add_argument("--cq", type=int, default=8)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    N, C, Cq = a.n, a.c, a.cq
    nq = 2 * Cq + C
    dev = "cuda"
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn(N, nq, generator=g).to(dev).half()
    qkvb = qkv.bfloat16()
    do = (torch.randn(N, C, generator=g) * 1e-2).to(dev).bfloat16()
    P = torch.empty(1, N, N, dtype=torch.float16, device=dev)
    Pb = torch.empty(1, N, N, dtype=torch.bfloat16, device=dev)
    o = torch.empty(1, N, C, device=dev)
    dqkv = torch.empty(N, nq, device=dev)
    D = torch.randn(N, generator=g).to(dev) * 1e-3
    parts = L.lib().dfcsa_bgemm_rowstat_parts(N)
    rowstat = torch.empty(1, parts, N, 2, device=dev)
    lse = torch.empty(N, device=dev)
    q, k, v = qkv[:, :Cq], qkv[:, Cq:2 * Cq], qkv[:, 2 * Cq:]
    qb, kb, vb = qkvb[:, :Cq], qkvb[:, Cq:2 * Cq], qkvb[:, 2 * Cq:]
    steps = [
        ("rowstats   (no output)", 0.0, lambda: ops.bgemm(1, N, N, Cq, q, N * nq, nq, False, k, N * nq, nq, False, None, 0, 0, epi=1, rowstat=rowstat)),
        ("lse_combine", parts * N * 8.0, lambda: L.call("dfcsa_lse_combine", L.ptr(rowstat), parts, 1, N, L.ptr(lse), L.stream())),
        ("exp -> P fp16 (write N^2)", 2.0 * N * N, lambda: ops.bgemm(1, N, N, Cq, q, N * nq, nq, False, k, N * nq, nq, False, P, N * N, N, epi=2, rowvec=lse)),
        ("o = P v (read N^2, K-major A)", 2.0 * N * N, lambda: ops.bgemm(1, N, C, N, P, N * N, N, False, v, N * nq, nq, True, o, N * C, C)),
        ("fused exp + P v (no N^2 traffic)", 0.0, lambda: ops.attn_pv_fused(qkv, 1, N, Cq, C, lse, o[0:1])),
        ("exp -> P bf16 (write N^2)", 2.0 * N * N, lambda: ops.bgemm(1, N, N, Cq, q, N * nq, nq, False, k, N * nq, nq, False, Pb, N * N, N, epi=2, rowvec=lse)),
        ("dv = P^T dO (read N^2, MN-major A)", 2.0 * N * N, lambda: ops.bgemm(1, N, C, N, Pb, N * N, N, True, do, N * C, C, True, dqkv[:, 2 * Cq:], N * nq, nq)),
        ("fused backward dq | dk | dv (2 kernels)", 0.0, lambda: ops.attn_bwd_fused(qkv, qkvb, do, 1, N, Cq, C, lse, D, dqkv)),
        ("dS = P*(dO v^T - D) in place (r+w N^2)", 4.0 * N * N, lambda: ops.softmax_bwd_bgemm(1, N, N, C, do, N * C, C, vb, N * nq, nq, Pb, D, Pb)),
        ("dq = dS k (read N^2, K-major A)", 2.0 * N * N, lambda: ops.bgemm(1, N, Cq, N, Pb, N * N, N, False, kb, N * nq, nq, True, dqkv[:, :Cq], N * nq, nq)),
        ("dk = dS^T q (read N^2, MN-major A)", 2.0 * N * N, lambda: ops.bgemm(1, N, Cq, N, Pb, N * N, N, True, qb, N * nq, nq, True, dqkv[:, Cq:2 * Cq], N * nq, nq)),
    ]
    for name, nbytes, fn in steps:
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        print(f"{name:42s} {ms:8.3f} ms   {nbytes / ms / 1e9 if nbytes else 0:7.2f} TB/s", flush=True)


if __name__ == "__main__":
    main()
