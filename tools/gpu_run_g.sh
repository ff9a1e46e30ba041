#!/bin/bash
# round-2 GPU session G (1 GPU): tests, headline bench, C4 at N = 1, ncu launch list of two eager steps, ncu of the conv GEMMs
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_ddp.py > gpurun_out/g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/g_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --detail gpurun_out/g_shapes.json > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
timeout 900 python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/g_c4_n1.json 2> gpurun_out/g_c4_n1.err
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-reference-gpu"
timeout 600 $CMD > gpurun_out/g_plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1100 --csv --log-file gpurun_out/g_launches.csv $CMD > gpurun_out/g_ncu_launches.log 2>&1
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -k regex:conv_tc -s 258 -c 86 --csv --log-file gpurun_out/g_ncu_conv_tc_traffic.csv $CMD > gpurun_out/g_ncu_traffic.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 258 -c 12 -f -o gpurun_out/g_ncu_conv_tc_full $CMD > gpurun_out/g_ncu_full.log 2>&1
ls -la gpurun_out/g_*
tail -n 4 gpurun_out/g_tests.log
head -c 300 gpurun_out/g_bench.json; echo
head -c 300 gpurun_out/g_c4_n1.json
