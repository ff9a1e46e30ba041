#!/bin/bash
# round-2 GPU session Y (1 GPU): the image as 64-wide patch rows (first block on the tensor-core path) - tests, A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "patch3x3" > gpurun_out/y_tests_patch.log 2>&1; echo "tests rc=$?" >> gpurun_out/y_tests_patch.log
tail -n 3 gpurun_out/y_tests_patch.log
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q -x > gpurun_out/y_tests_net.log 2>&1; echo "tests rc=$?" >> gpurun_out/y_tests_net.log
tail -n 12 gpurun_out/y_tests_net.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu"
timeout 600 $B > gpurun_out/y_bench_on.json 2> gpurun_out/y_bench_on.err
DFCSA_PATCH_INPUT=0 timeout 600 $B > gpurun_out/y_bench_off.json 2> gpurun_out/y_bench_off.err
timeout 600 $B > gpurun_out/y_bench_on2.json 2> gpurun_out/y_bench_on2.err
timeout 600 python tools/bench_configs.py c5 --out gpurun_out/y_configs_on.json > gpurun_out/y_configs_on.log 2>&1
for f in on off on2; do head -c 200 gpurun_out/y_bench_$f.json; echo; tail -n 3 gpurun_out/y_bench_$f.err; done
grep -E "^c5.*(b8|b1)_eval" gpurun_out/y_configs_on.log | cut -c1-150
