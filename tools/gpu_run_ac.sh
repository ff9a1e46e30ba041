#!/bin/bash
# round-2 GPU session AC (1 GPU): where the pool-32 configuration loses its time - per-kernel block tables at P = 32 and P = 16
mkdir -p gpurun_out
for lv in 1 3 5; do
  timeout 200 python tools/block_bench.py --level $lv --pool 32 --out gpurun_out/ac_p32_l$lv.json > gpurun_out/ac_p32_l$lv.log 2>&1
  timeout 200 python tools/block_bench.py --level $lv --pool 16 --out gpurun_out/ac_p16_l$lv.json > gpurun_out/ac_p16_l$lv.log 2>&1
done
tail -n 30 gpurun_out/ac_p32_l5.log
