#!/bin/bash
# round-2 GPU session X (1 GPU): final tree after the elect.sync change - all tests, default bench line, ncu launch list, conv traffic, ncu --set full of the K = 9216 pair GEMM, all configs
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_ddp.py > gpurun_out/x_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/x_tests.log
timeout 900 python bench.py > gpurun_out/x_bench_default.json 2> gpurun_out/x_bench_default.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --detail gpurun_out/x_shapes.json > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-reference-gpu"
timeout 600 $CMD > gpurun_out/x_plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 1000 --csv --log-file gpurun_out/x_launches.csv $CMD > gpurun_out/x_ncu_launches.log 2>&1
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -k regex:conv_tc -s 258 -c 86 --csv --log-file gpurun_out/x_ncu_conv_tc_traffic.csv $CMD > gpurun_out/x_ncu_traffic.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 282 -c 3 -f -o gpurun_out/x_ncu_conv_tc_pairs_full $CMD > gpurun_out/x_ncu_full.log 2>&1
timeout 900 python tools/bench_configs.py c1 c2 c3 c5 --out gpurun_out/x_configs.json > gpurun_out/x_configs.log 2>&1
ls -la gpurun_out/x_*
tail -n 4 gpurun_out/x_tests.log
head -c 300 gpurun_out/x_bench_default.json; echo
head -c 300 gpurun_out/x_bench.json; echo
grep -E "^c[0-9]" gpurun_out/x_configs.log | cut -c1-160
