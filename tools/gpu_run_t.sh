#!/bin/bash
# round-2 GPU session T (1 GPU): order of the tcgen05.mma instructions in the 3x3 weight gradient (same accumulator back to back vs round-robin over the three tap accumulators)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k wgrad > gpurun_out/t_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_tests.log
tail -n 3 gpurun_out/t_tests.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu"
DFCSA_WGRAD_ORDER=0 timeout 600 $B --detail gpurun_out/t_shapes_0.json > gpurun_out/t_bench_0.json 2> gpurun_out/t_bench_0.err
DFCSA_WGRAD_ORDER=1 timeout 600 $B --detail gpurun_out/t_shapes_1.json > gpurun_out/t_bench_1.json 2> gpurun_out/t_bench_1.err
for f in 0 1; do head -c 200 gpurun_out/t_bench_$f.json; echo; tail -n 2 gpurun_out/t_bench_$f.err; done
