#!/bin/bash
# round-2 GPU session F (1 GPU): fast sigmoid + two pixels in flight in the gate kernels, pool_rows with four loads in flight
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_ddp.py > gpurun_out/f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/f_tests.log
for lvl in 1 2; do
  timeout 300 python tools/block_bench.py --level $lvl --out gpurun_out/f_block${lvl}.json > gpurun_out/f_block${lvl}.txt 2>&1
  DFCSA_EW_OCC=2 timeout 300 python tools/block_bench.py --level $lvl --out gpurun_out/f_block${lvl}_occ2.json > gpurun_out/f_block${lvl}_occ2.txt 2>&1
  DFCSA_EW_OCC=4 timeout 300 python tools/block_bench.py --level $lvl --out gpurun_out/f_block${lvl}_occ4.json > gpurun_out/f_block${lvl}_occ4.txt 2>&1
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/f_bench2.json 2> gpurun_out/f_bench2.err
tail -n 4 gpurun_out/f_tests.log
grep -E "gate_mix|bnrelu_pool|block Ci" gpurun_out/f_block1.txt gpurun_out/f_block1_occ2.txt gpurun_out/f_block1_occ4.txt
head -c 300 gpurun_out/f_bench.json
