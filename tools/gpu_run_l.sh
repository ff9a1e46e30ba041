#!/bin/bash
# round-2 GPU session L (1 GPU): CTA-pair (cta_group::2) conv GEMM: tests first, then A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "conv_gemm" > gpurun_out/l_conv_tests.log 2>&1; rc=$?; echo "conv tests rc=$rc" >> gpurun_out/l_conv_tests.log
tail -n 12 gpurun_out/l_conv_tests.log
if [ "$rc" != "0" ]; then
  DFCSA_CONV_2CTA=0 timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "conv_gemm" > gpurun_out/l_conv_tests_1cta.log 2>&1; echo "1cta rc=$?" >> gpurun_out/l_conv_tests_1cta.log
  tail -n 5 gpurun_out/l_conv_tests_1cta.log
  exit 0
fi
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_ddp.py > gpurun_out/l_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/l_tests.log
tail -n 6 gpurun_out/l_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --detail gpurun_out/l_shapes_2cta.json > gpurun_out/l_bench_2cta.json 2> gpurun_out/l_bench_2cta.err
DFCSA_CONV_2CTA=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --detail gpurun_out/l_shapes_1cta.json > gpurun_out/l_bench_1cta.json 2> gpurun_out/l_bench_1cta.err
head -c 260 gpurun_out/l_bench_2cta.json; echo; head -c 260 gpurun_out/l_bench_1cta.json; echo
tail -n 3 gpurun_out/l_bench_2cta.err
