#!/bin/bash
# round-2 GPU session O (1 GPU): weight gradients on a second stream - parity tests, A/B of the step time (off / same priority / higher priority)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q -x > gpurun_out/o_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/o_tests.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu"
DFCSA_WGRAD_STREAM=0 timeout 600 $B > gpurun_out/o_bench_off.json 2> gpurun_out/o_bench_off.err
DFCSA_WGRAD_STREAM_PRIO=0 timeout 600 $B > gpurun_out/o_bench_p0.json 2> gpurun_out/o_bench_p0.err
DFCSA_WGRAD_STREAM_PRIO=-1 timeout 600 $B > gpurun_out/o_bench_pm1.json 2> gpurun_out/o_bench_pm1.err
DFCSA_WGRAD_STREAM=0 timeout 600 $B > gpurun_out/o_bench_off2.json 2> gpurun_out/o_bench_off2.err
DFCSA_WGRAD_STREAM_PRIO=0 timeout 600 $B --no-graph > gpurun_out/o_bench_p0_eager.json 2> gpurun_out/o_bench_p0_eager.err
timeout 600 python tools/bench_configs.py c1 --out gpurun_out/o_configs.json > gpurun_out/o_configs.log 2>&1
tail -n 4 gpurun_out/o_tests.log
for f in off p0 pm1 off2 p0_eager; do echo $f; head -c 200 gpurun_out/o_bench_$f.json; echo; tail -n 3 gpurun_out/o_bench_$f.err; done
grep -E "^c[0-9]" gpurun_out/o_configs.log | cut -c1-160
