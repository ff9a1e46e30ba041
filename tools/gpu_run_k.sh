#!/bin/bash
# timing experiment: the weight gradient without the fp16 -> bf16 rewrite (numerically wrong on purpose) = upper bound of what fp16 gradients would buy
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --detail gpurun_out/k_shapes_cvt.json > gpurun_out/k_bench_cvt.json 2> gpurun_out/k_bench_cvt.err
DFCSA_WGRAD_NOCVT_TIMING_ONLY=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --detail gpurun_out/k_shapes_nocvt.json > gpurun_out/k_bench_nocvt.json 2> gpurun_out/k_bench_nocvt.err
head -c 250 gpurun_out/k_bench_cvt.json; echo; head -c 250 gpurun_out/k_bench_nocvt.json; echo
tail -3 gpurun_out/k_bench_nocvt.err
