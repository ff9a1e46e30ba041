#!/bin/bash
# round-2 GPU session V (1 GPU): CTA-pair K threshold again now that the issue path is short; e2e after its warm-up fix
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu"
for kb in 16 8 4 2; do
  DFCSA_CONV_2CTA_MINKB=$kb timeout 600 $B --detail gpurun_out/v_shapes_kb$kb.json > gpurun_out/v_bench_kb$kb.json 2> gpurun_out/v_bench_kb$kb.err
  head -c 200 gpurun_out/v_bench_kb$kb.json; echo; tail -n 2 gpurun_out/v_bench_kb$kb.err
done
DFCSA_CONV_2CTA=0 timeout 600 $B --detail gpurun_out/v_shapes_nopairs.json > gpurun_out/v_bench_nopairs.json 2> gpurun_out/v_bench_nopairs.err
