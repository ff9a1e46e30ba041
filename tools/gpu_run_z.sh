#!/bin/bash
# round-2 GPU session Z (1 GPU): ncu --set full with source counters AFTER the elect.sync change: level-1 decoder 3x3 conv (N=64, K=1152) and the level-1 3x3 weight gradient
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-reference-gpu"
timeout 600 $CMD > gpurun_out/z_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 300 -c 1 -f -o gpurun_out/z_ncu_conv_n64_k1152 $CMD > gpurun_out/z_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc -s 57 -c 1 -f -o gpurun_out/z_ncu_wgrad_l1_3x3 $CMD > gpurun_out/z_ncu2.log 2>&1
tail -n 2 gpurun_out/z_ncu1.log gpurun_out/z_ncu2.log
ls -la gpurun_out/z_*
