#!/bin/bash
# round-2 GPU session W (1 GPU): resident weights again, on top of the short issue path
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "test_conv_gemm or epilogue" > gpurun_out/w_tests_conv.log 2>&1; echo "tests rc=$?" >> gpurun_out/w_tests_conv.log
tail -n 3 gpurun_out/w_tests_conv.log
if ! grep -q "rc=0" gpurun_out/w_tests_conv.log; then exit 1; fi
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu"
timeout 600 $B --detail gpurun_out/w_shapes_on.json > gpurun_out/w_bench_on.json 2> gpurun_out/w_bench_on.err
DFCSA_CONV_BRES=0 timeout 600 $B --detail gpurun_out/w_shapes_off.json > gpurun_out/w_bench_off.json 2> gpurun_out/w_bench_off.err
timeout 600 $B > gpurun_out/w_bench_on2.json 2> gpurun_out/w_bench_on2.err
timeout 600 python tools/bench_configs.py c5 --out gpurun_out/w_configs_on.json > gpurun_out/w_configs_on.log 2>&1
DFCSA_CONV_BRES=0 timeout 600 python tools/bench_configs.py c5 --out gpurun_out/w_configs_off.json > gpurun_out/w_configs_off.log 2>&1
for f in on off on2; do head -c 200 gpurun_out/w_bench_$f.json; echo; done
grep -E "^c5.*(b8|b1)_eval" gpurun_out/w_configs_on.log gpurun_out/w_configs_off.log | cut -c1-150
