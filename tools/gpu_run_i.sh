#!/bin/bash
# round-2 GPU session I (1 GPU): BatchNorm finalize folded into the GEMM (A/B), C1, per-kernel times at batch 4, default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_ddp.py > gpurun_out/i_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/i_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/i_bench_fold.json 2> gpurun_out/i_bench_fold.err
DFCSA_BN_FOLD=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/i_bench_nofold.json 2> gpurun_out/i_bench_nofold.err
timeout 600 python tools/bench_configs.py c1 --profile --out gpurun_out/i_c1_fold.json > gpurun_out/i_c1_fold.log 2>&1
DFCSA_BN_FOLD=0 timeout 600 python tools/bench_configs.py c1 --out gpurun_out/i_c1_nofold.json > gpurun_out/i_c1_nofold.log 2>&1
for lvl in 1 3 5; do timeout 200 python tools/block_bench.py --level $lvl --batch 4 --steps 20 > gpurun_out/i_block${lvl}_b4.txt 2>&1; done
timeout 900 python bench.py > gpurun_out/i_bench_default.json 2> gpurun_out/i_bench_default.err
tail -n 4 gpurun_out/i_tests.log
head -c 260 gpurun_out/i_bench_fold.json; echo
head -c 260 gpurun_out/i_bench_nofold.json; echo
grep -E "^c1" gpurun_out/i_c1_fold.log gpurun_out/i_c1_nofold.log | cut -c1-200
head -c 400 gpurun_out/i_bench_default.json
