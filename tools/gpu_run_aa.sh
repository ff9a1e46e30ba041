#!/bin/bash
# round-2 GPU session AA (1 GPU): two pixels in flight in branch_bwd_reduce1 - tests, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_streaming.py -m gpu -q -x -k "branch or streaming or reduce" > gpurun_out/aa_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/aa_tests.log
tail -n 3 gpurun_out/aa_tests.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu"
timeout 600 $B > gpurun_out/aa_bench.json 2> gpurun_out/aa_bench.err
timeout 600 $B > gpurun_out/aa_bench2.json 2> gpurun_out/aa_bench2.err
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q -x -k "full_width or trajectory or dice" > gpurun_out/aa_tests_net.log 2>&1; echo "tests rc=$?" >> gpurun_out/aa_tests_net.log
tail -n 3 gpurun_out/aa_tests_net.log
for f in bench bench2; do head -c 200 gpurun_out/aa_$f.json; echo; done
