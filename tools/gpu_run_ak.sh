#!/bin/bash
# round-2 GPU session AK (1 GPU): look-ahead loads in pool_rows / bilerpT_rows - all single-GPU tests, per-kernel A/B on a level-1 / level-2 block, step A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_ddp.py > gpurun_out/ak_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/ak_tests.log
tail -n 3 gpurun_out/ak_tests.log
BB="python tools/block_bench.py --steps 5 --warmup 2"
timeout 120 $BB --level 1 --out gpurun_out/ak_block1_default.json > gpurun_out/ak_block1_default.log 2>&1
DFCSA_POOL_PF=0 DFCSA_BILERPT_PF=0 timeout 120 $BB --level 1 --out gpurun_out/ak_block1_nopf.json > gpurun_out/ak_block1_nopf.log 2>&1
timeout 120 $BB --level 2 --out gpurun_out/ak_block2_default.json > gpurun_out/ak_block2_default.log 2>&1
DFCSA_POOL_PF=0 DFCSA_BILERPT_PF=0 timeout 120 $BB --level 2 --out gpurun_out/ak_block2_nopf.json > gpurun_out/ak_block2_nopf.log 2>&1
grep -H "branch_bwd_reduce1\|bnrelu_pool_fwd" gpurun_out/ak_block*.log
B="python bench.py --warmup 3 --no-cpu-baseline --no-reference-gpu --steps 10"
timeout 300 $B > gpurun_out/ak_bench_default.json 2> gpurun_out/ak_bench_default.err
DFCSA_POOL_PF=0 DFCSA_BILERPT_PF=0 timeout 300 $B > gpurun_out/ak_bench_nopf.json 2> gpurun_out/ak_bench_nopf.err
for f in default nopf; do head -c 200 gpurun_out/ak_bench_$f.json; echo; tail -n 2 gpurun_out/ak_bench_$f.err; done
