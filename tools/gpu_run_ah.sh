#!/bin/bash
# round-2 GPU session AH (1 GPU): branch_bwd_reduce1 with two pixels in flight, pool_rows with eight - kernel tests, per-kernel
# A/B on one level-1 / level-2 block, step A/B, and ncu --set full of the level-1 weight gradients and of reduce1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_streaming.py tests/test_gpu_conv.py -m gpu -q -x -k "branch or pool or window" > gpurun_out/ah_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/ah_tests.log
tail -n 3 gpurun_out/ah_tests.log
BB="python tools/block_bench.py --steps 5 --warmup 2"
timeout 120 $BB --level 1 --out gpurun_out/ah_block1_default.json > gpurun_out/ah_block1_default.log 2>&1
DFCSA_RED1_PIX=1 timeout 120 $BB --level 1 --out gpurun_out/ah_block1_red1pix1.json > gpurun_out/ah_block1_red1pix1.log 2>&1
DFCSA_POOL_PIX=8 timeout 120 $BB --level 1 --out gpurun_out/ah_block1_poolpix8.json > gpurun_out/ah_block1_poolpix8.log 2>&1
timeout 120 $BB --level 2 --out gpurun_out/ah_block2_default.json > gpurun_out/ah_block2_default.log 2>&1
DFCSA_RED1_PIX=1 DFCSA_POOL_PIX=8 timeout 120 $BB --level 2 --out gpurun_out/ah_block2_red1pix1_poolpix8.json > gpurun_out/ah_block2_red1pix1_poolpix8.log 2>&1
grep -h "branch_bwd_reduce1\|bnrelu_pool_fwd" gpurun_out/ah_block*.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu"
timeout 300 $B > gpurun_out/ah_bench_default.json 2> gpurun_out/ah_bench_default.err
DFCSA_RED1_PIX=1 timeout 300 $B > gpurun_out/ah_bench_red1pix1.json 2> gpurun_out/ah_bench_red1pix1.err
DFCSA_POOL_PIX=8 timeout 300 $B > gpurun_out/ah_bench_poolpix8.json 2> gpurun_out/ah_bench_poolpix8.err
for f in default red1pix1 poolpix8; do head -c 200 gpurun_out/ah_bench_$f.json; echo; done
NB="python tools/block_bench.py --steps 1 --warmup 1 --level 1"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc -c 6 -f -o gpurun_out/ah_ncu_wgrad_l1 $NB > gpurun_out/ah_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:branch_bwd_reduce1 -c 2 -f -o gpurun_out/ah_ncu_reduce1_l1 $NB > gpurun_out/ah_ncu2.log 2>&1
tail -n 2 gpurun_out/ah_ncu1.log gpurun_out/ah_ncu2.log
ls -la gpurun_out/ah_*
