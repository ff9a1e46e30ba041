#!/bin/bash
# round-2 GPU session AF (1 GPU): programmatic dependent launch on every kernel - all tests, step and C1 / C5 with and without
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_ddp.py > gpurun_out/af_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/af_tests.log
tail -n 4 gpurun_out/af_tests.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu"
timeout 300 $B > gpurun_out/af_bench_on.json 2> gpurun_out/af_bench_on.err
DFCSA_PDL=0 timeout 300 $B > gpurun_out/af_bench_off.json 2> gpurun_out/af_bench_off.err
timeout 300 $B > gpurun_out/af_bench_on2.json 2> gpurun_out/af_bench_on2.err
timeout 300 python tools/bench_configs.py c1 c5 --out gpurun_out/af_configs_on.json > gpurun_out/af_configs_on.log 2>&1
DFCSA_PDL=0 timeout 300 python tools/bench_configs.py c1 c5 --out gpurun_out/af_configs_off.json > gpurun_out/af_configs_off.log 2>&1
for f in on off on2; do head -c 200 gpurun_out/af_bench_$f.json; echo; tail -n 2 gpurun_out/af_bench_$f.err; done
grep -E "^c1|^c5.*(b1|b8)_eval" gpurun_out/af_configs_on.log | cut -c1-140
grep -E "^c1|^c5.*(b1|b8)_eval" gpurun_out/af_configs_off.log | cut -c1-140
