#!/usr/bin/env python
"""Per-kernel timings of ONE DFC-SA block (forward + backward) at a chosen level of the network, for kernel work that
does not need a whole training step: a level-1 block at batch 64 takes ~10 ms, so a dozen variants of one streaming
kernel (DFCSA_EW_OCC, DFCSA_ACT_OCC, DFCSA_WGRAD_WAVES, ... - one process per setting) fit in one short GPU call.

  python tools/block_bench.py --level 1 --batch 64            # Ci=128 -> Co=64 at 224^2 (the decoder's last block)
  python tools/block_bench.py --ci 64 --co 128 --hw 112       # explicit shape

Every libdfcsa entry point is timed with a CUDA-event pair (dfcsa._lib.Profiler, the same instrument as bench.py's
`kernels` table).  For the streaming kernels the table adds effective GB/s from the algorithmic element moves per pass
(E = batch * Co * hw^2 elements of 2 bytes; the per-kernel E counts are the ones in profiles/README.md) and the fraction
of the measured copy bandwidth (MEASURED_PEAKS.json, else 6459 GB/s)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dfc-sa-unet_b200")]

import torch  # noqa: E402

from dfcsa import _lib  # noqa: E402
from dfcsa.modules import DynamicFusionConvAttnBlock  # noqa: E402

# algorithmic element moves (reads + writes) per pass, in units of E
E_PER_PASS = {
    "bnrelu_pool_fwd": 1.0, "branch_act_fwd": 4.0, "gate_mix_fwd": 4.0, "block_out_fwd": 3.0,
    "block_out_bwd_reduce": 4.0, "bn_bwd_apply": 3.0, "gate_mix_bwd_reduce": 4.0, "gate_mix_bwd_apply": 5.0,
    "branch_bwd_reduce1": 8.0, "branch_bwd_reduce2": 2.0, "branch_bwd_apply": 6.0,
}
# decoder-side block shapes of the 224^2 network (Ci, Co, hw): level 1 .. 4, and the bottleneck as level 5
LEVELS = {1: (128, 64, 224), 2: (256, 128, 112), 3: (512, 256, 56), 4: (1024, 512, 28), 5: (512, 1024, 14)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=1, choices=sorted(LEVELS))
    ap.add_argument("--ci", type=int, default=None)
    ap.add_argument("--co", type=int, default=None)
    ap.add_argument("--hw", type=int, default=None)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--pool", type=int, default=4)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--out", default=None, help="also write the table as JSON")
    args = ap.parse_args()
    ci, co, hw = LEVELS[args.level]
    ci, co, hw = args.ci or ci, args.co or co, args.hw or hw
    if not torch.cuda.is_available():
        raise SystemExit("block_bench needs a CUDA device")
    peak = 6459.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except (OSError, KeyError, ValueError):
        pass
    torch.manual_seed(0)
    blk = DynamicFusionConvAttnBlock(ci, co, pool_size=args.pool, ablation_on_qk_channels=8).cuda().train()
    with torch.no_grad():
        blk.attn_branch[3].gamma.fill_(0.5)
    x = torch.randn(args.batch, ci, hw, hw, device="cuda", requires_grad=True)
    r = torch.randn(args.batch, co, hw, hw, device="cuda")

    def step():
        for p in blk.parameters():
            p.grad = None
        x.grad = None
        y = blk(x)
        y.backward(r)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    prof = _lib.Profiler()
    _lib.PROF = prof
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    _lib.PROF = None
    total_ms = e0.elapsed_time(e1) / args.steps
    e_bytes = 2.0 * args.batch * co * hw * hw
    rows = []
    for tag, a in sorted(prof.summary().items(), key=lambda kv: -kv[1]["ms"]):
        ms = a["ms"] / args.steps
        row = {"kernel": tag, "ms": round(ms, 4), "launches": a["launches"] / args.steps}
        if a["flops"] > 0:
            row["tflops"] = round(a["flops"] / args.steps / (ms * 1e-3) / 1e12, 1)
        if tag in E_PER_PASS:
            gbs = E_PER_PASS[tag] * e_bytes / (ms * 1e-3) / 1e9
            row["gb_s"] = round(gbs, 0)
            row["of_copy_bw"] = round(gbs / peak, 2)
        rows.append(row)
    print(f"block Ci={ci} Co={co} {hw}x{hw} batch {args.batch} P={args.pool}: {total_ms:.3f} ms per forward+backward "
          f"(eager, instrumented); E = {e_bytes / 1e6:.0f} MB; copy bandwidth {peak:.0f} GB/s")
    for row in rows:
        print("  " + "  ".join(f"{k}={v}" for k, v in row.items()))
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)) or ".", exist_ok=True)
        json.dump({"shape": {"ci": ci, "co": co, "hw": hw, "batch": args.batch, "pool": args.pool}, "ms_total": total_ms,
                   "kernels": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
