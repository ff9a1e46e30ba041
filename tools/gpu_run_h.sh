#!/bin/bash
# round-2 GPU session H (1 GPU): trimmed conv epilogue + pool rows; the other BASELINE configurations on the final tree
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_ddp.py > gpurun_out/h_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/h_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --detail gpurun_out/h_shapes.json > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/h_bench2.json 2> gpurun_out/h_bench2.err
timeout 1500 python tools/bench_configs.py c1 c2 c3 c5 --profile --out gpurun_out/h_configs.json > gpurun_out/h_configs.log 2>&1
timeout 300 python tools/block_bench.py --level 1 --out gpurun_out/h_block1.json > gpurun_out/h_block1.txt 2>&1
tail -n 4 gpurun_out/h_tests.log
head -c 300 gpurun_out/h_bench.json; echo
head -c 300 gpurun_out/h_bench2.json; echo
grep -E "^c[0-9]" gpurun_out/h_configs.log | cut -c1-220
