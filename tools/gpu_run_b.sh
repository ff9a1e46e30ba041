#!/bin/bash
# round-2 GPU session B: arena / one-launch parameter gradients / folded inference path; C1 and C5; reference-on-GPU profile
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_ddp.py > gpurun_out/b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/b_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --detail gpurun_out/b_shapes.json > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
timeout 900 python tools/bench_configs.py c1 c5 --out gpurun_out/b_configs.json > gpurun_out/b_configs.log 2>&1
timeout 600 python tools/ref_gpu_profile.py > gpurun_out/b_ref_gpu_profile.txt 2>&1
tail -n 5 gpurun_out/b_tests.log
head -c 400 gpurun_out/b_bench.json
