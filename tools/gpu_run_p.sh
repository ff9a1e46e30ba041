#!/bin/bash
# round-2 GPU session P (1 GPU): fused inference epilogues (gate mix / residual in the GEMM) - tests, C5 A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_net.py -m gpu -q > gpurun_out/p_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/p_tests.log
DFCSA_EVAL_EPI=0 timeout 600 python tools/bench_configs.py c5 --out gpurun_out/p_configs_off.json > gpurun_out/p_configs_off.log 2>&1
timeout 600 python tools/bench_configs.py c5 --profile --out gpurun_out/p_configs_on.json > gpurun_out/p_configs_on.log 2>&1
tail -n 5 gpurun_out/p_tests.log
grep -E "^c[0-9]" gpurun_out/p_configs_off.log | cut -c1-160
grep -E "^c[0-9]" gpurun_out/p_configs_on.log | cut -c1-160
