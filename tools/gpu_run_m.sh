#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "conv_gemm" > gpurun_out/m_conv_tests.log 2>&1; echo "conv tests rc=$?" >> gpurun_out/m_conv_tests.log
tail -n 30 gpurun_out/m_conv_tests.log | cut -c1-220
timeout 300 python -m pytest tests/test_gpu_net.py -m gpu -q -x -k "baseline_configs and 224-4-4" > gpurun_out/m_net_test.log 2>&1; echo "net rc=$?" >> gpurun_out/m_net_test.log
grep -E "mbarrier|passed|failed" gpurun_out/m_net_test.log | sort | uniq -c | sort -rn | head -12
