#!/bin/bash
# round-2 GPU session R (1 GPU): incremental UMMA descriptors in the MMA issue loops (conv_tc, wgrad_tc) - tests, step A/B with / without resident weights
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q -x > gpurun_out/r_tests_conv.log 2>&1; echo "tests rc=$?" >> gpurun_out/r_tests_conv.log
tail -n 3 gpurun_out/r_tests_conv.log
if ! grep -q "rc=0" gpurun_out/r_tests_conv.log; then exit 1; fi
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu"
timeout 600 $B --detail gpurun_out/r_shapes_on.json > gpurun_out/r_bench_on.json 2> gpurun_out/r_bench_on.err
DFCSA_CONV_BRES=0 timeout 600 $B --detail gpurun_out/r_shapes_off.json > gpurun_out/r_bench_off.json 2> gpurun_out/r_bench_off.err
timeout 600 python tools/bench_configs.py c1 c5 --out gpurun_out/r_configs.json > gpurun_out/r_configs.log 2>&1
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q > gpurun_out/r_tests_net.log 2>&1; echo "tests rc=$?" >> gpurun_out/r_tests_net.log
tail -n 3 gpurun_out/r_tests_net.log
for f in off on; do head -c 200 gpurun_out/r_bench_$f.json; echo; tail -n 2 gpurun_out/r_bench_$f.err; done
grep -E "^c[0-9]" gpurun_out/r_configs.log | cut -c1-130
