#!/usr/bin/env python
"""Where the UNMODIFIED reference spends its time in eager PyTorch on this GPU (bench.py --impl reference-gpu measures 71 img/s
at batch 64): torch.profiler over one training step, top CUDA kernels by total time.

  python tools/ref_gpu_profile.py [--batch 64] [--autocast] > gpurun_out/ref_gpu_profile.txt
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dfc-sa-unet_b200")]

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--autocast", action="store_true")
    a = ap.parse_args()
    from dfcsa.synthetic import synthetic_batch
    imgs, masks = zip(*[synthetic_batch(16, 224, 224, seed=10 + j) for j in range((a.batch + 15) // 16)])
    img, mask = torch.cat(imgs)[:a.batch].cuda(), torch.cat(masks)[:a.batch].cuda()
    step = bench.reference_step_fn("cuda", autocast_bf16=a.autocast)
    for _ in range(2):
        step(img, mask)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step(img, mask)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=90))


if __name__ == "__main__":
    main()
