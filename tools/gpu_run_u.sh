#!/bin/bash
# round-2 GPU session U (1 GPU): tcgen05.mma / TMA issue under elect.sync (no per-instruction ELECT / BRA.U.ANY wrapper) - tests, step, configs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q -x > gpurun_out/u_tests_conv.log 2>&1; echo "tests rc=$?" >> gpurun_out/u_tests_conv.log
tail -n 3 gpurun_out/u_tests_conv.log
if ! grep -q "rc=0" gpurun_out/u_tests_conv.log; then exit 1; fi
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu"
timeout 600 $B --detail gpurun_out/u_shapes_1.json > gpurun_out/u_bench_1.json 2> gpurun_out/u_bench_1.err
DFCSA_WGRAD_ORDER=0 timeout 600 $B --detail gpurun_out/u_shapes_0.json > gpurun_out/u_bench_0.json 2> gpurun_out/u_bench_0.err
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q > gpurun_out/u_tests_net.log 2>&1; echo "tests rc=$?" >> gpurun_out/u_tests_net.log
tail -n 3 gpurun_out/u_tests_net.log
timeout 600 python tools/bench_configs.py c1 c5 --out gpurun_out/u_configs.json > gpurun_out/u_configs.log 2>&1
for f in 0 1; do head -c 200 gpurun_out/u_bench_$f.json; echo; tail -n 2 gpurun_out/u_bench_$f.err; done
grep -E "^c[0-9]" gpurun_out/u_configs.log | cut -c1-130
