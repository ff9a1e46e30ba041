#!/bin/bash
# round-2 GPU session AI (1 GPU): one-cell bilinear^T rows, CTA-per-cell column reductions, preloading max-pool backward
# reduction - kernel tests, per-kernel A/B on one level-1 / level-2 block, step A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_streaming.py -m gpu -q -x > gpurun_out/ai_tests_streaming.log 2>&1; echo "tests rc=$?" >> gpurun_out/ai_tests_streaming.log
tail -n 3 gpurun_out/ai_tests_streaming.log
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "branch or pool or block_out or attention" > gpurun_out/ai_tests_conv.log 2>&1; echo "tests rc=$?" >> gpurun_out/ai_tests_conv.log
tail -n 3 gpurun_out/ai_tests_conv.log
BB="python tools/block_bench.py --steps 5 --warmup 2"
timeout 120 $BB --level 1 --out gpurun_out/ai_block1_default.json > gpurun_out/ai_block1_default.log 2>&1
DFCSA_BILERPT_ONE_CELL=0 DFCSA_COLS_PAR=0 timeout 120 $BB --level 1 --out gpurun_out/ai_block1_old.json > gpurun_out/ai_block1_old.log 2>&1
timeout 120 $BB --level 2 --out gpurun_out/ai_block2_default.json > gpurun_out/ai_block2_default.log 2>&1
DFCSA_BILERPT_ONE_CELL=0 DFCSA_COLS_PAR=0 timeout 120 $BB --level 2 --out gpurun_out/ai_block2_old.json > gpurun_out/ai_block2_old.log 2>&1
grep -H "branch_bwd_reduce1\|bnrelu_pool_fwd" gpurun_out/ai_block*.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu"
timeout 300 $B > gpurun_out/ai_bench_default.json 2> gpurun_out/ai_bench_default.err
DFCSA_BOUT_PRE=0 timeout 300 $B > gpurun_out/ai_bench_boutpre0.json 2> gpurun_out/ai_bench_boutpre0.err
DFCSA_BILERPT_ONE_CELL=0 DFCSA_COLS_PAR=0 timeout 300 $B > gpurun_out/ai_bench_rowsold.json 2> gpurun_out/ai_bench_rowsold.err
for f in default boutpre0 rowsold; do head -c 200 gpurun_out/ai_bench_$f.json; echo; tail -n 2 gpurun_out/ai_bench_$f.err; done
