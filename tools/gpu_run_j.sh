#!/bin/bash
# round-2 GPU session J (1 GPU): row-walk branch_act / branch_bwd_apply kernels (A/B), eval path
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_ddp.py > gpurun_out/j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/j_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/j_bench_rows.json 2> gpurun_out/j_bench_rows.err
DFCSA_ACT_ROWS=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/j_bench_norows.json 2> gpurun_out/j_bench_norows.err
for lvl in 1 2 3; do
  timeout 200 python tools/block_bench.py --level $lvl > gpurun_out/j_block${lvl}_rows.txt 2>&1
  DFCSA_ACT_ROWS=0 timeout 200 python tools/block_bench.py --level $lvl > gpurun_out/j_block${lvl}_norows.txt 2>&1
done
timeout 900 python tools/bench_configs.py c5 --profile --out gpurun_out/j_c5_rows.json > gpurun_out/j_c5_rows.log 2>&1
tail -n 4 gpurun_out/j_tests.log
head -c 260 gpurun_out/j_bench_rows.json; echo
head -c 260 gpurun_out/j_bench_norows.json; echo
grep -E "branch_act|branch_bwd_apply" gpurun_out/j_block*_rows.txt gpurun_out/j_block*_norows.txt
grep -E "^c5" gpurun_out/j_c5_rows.log | cut -c1-200
