#!/bin/bash
# round-2 GPU session Q (1 GPU): resident weights in conv_tc - tests, A/B of the step and of C5
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "test_conv_gemm or epilogue" > gpurun_out/q_tests_conv.log 2>&1; echo "tests rc=$?" >> gpurun_out/q_tests_conv.log
tail -n 3 gpurun_out/q_tests_conv.log
if ! grep -q "rc=0" gpurun_out/q_tests_conv.log; then exit 1; fi
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu"
DFCSA_CONV_BRES=0 timeout 600 $B --detail gpurun_out/q_shapes_off.json > gpurun_out/q_bench_off.json 2> gpurun_out/q_bench_off.err
timeout 600 $B --detail gpurun_out/q_shapes_on.json > gpurun_out/q_bench_on.json 2> gpurun_out/q_bench_on.err
DFCSA_CONV_BRES=0 timeout 600 python tools/bench_configs.py c5 --out gpurun_out/q_configs_off.json > gpurun_out/q_configs_off.log 2>&1
timeout 600 python tools/bench_configs.py c5 --profile --out gpurun_out/q_configs_on.json > gpurun_out/q_configs_on.log 2>&1
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q > gpurun_out/q_tests_net.log 2>&1; echo "tests rc=$?" >> gpurun_out/q_tests_net.log
tail -n 3 gpurun_out/q_tests_net.log
for f in off on; do head -c 200 gpurun_out/q_bench_$f.json; echo; tail -n 2 gpurun_out/q_bench_$f.err; done
grep -E "^c[0-9]" gpurun_out/q_configs_off.log | cut -c1-130
grep -E "^c[0-9]" gpurun_out/q_configs_on.log | cut -c1-130
