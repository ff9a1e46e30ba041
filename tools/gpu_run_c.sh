#!/bin/bash
# round-2 GPU session C: window-terms reduction A/B, eval profile, ncu of one level-1 block
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_ddp.py > gpurun_out/c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c_tests.log
for lvl in 1 2 3; do
  timeout 300 python tools/block_bench.py --level $lvl --out gpurun_out/c_block${lvl}_new.json > gpurun_out/c_block${lvl}_new.txt 2>&1
  DFCSA_OLD_REDUCE2=1 timeout 300 python tools/block_bench.py --level $lvl --out gpurun_out/c_block${lvl}_old.json > gpurun_out/c_block${lvl}_old.txt 2>&1
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/c_bench_new.json 2> gpurun_out/c_bench_new.err
DFCSA_OLD_REDUCE2=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/c_bench_old.json 2> gpurun_out/c_bench_old.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/c_bench_new2.json 2> gpurun_out/c_bench_new2.err
timeout 900 python tools/bench_configs.py c5 --profile --out gpurun_out/c_configs.json > gpurun_out/c_configs.log 2>&1
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section SchedulerStats --section Occupancy --section LaunchStats \
  --clock-control none -f -o gpurun_out/c_ncu_block1 python tools/block_bench.py --level 1 --steps 1 --warmup 1 > gpurun_out/c_ncu_block1.log 2>&1
ls -la gpurun_out/c_ncu_block1.ncu-rep
tail -n 4 gpurun_out/c_tests.log
