#!/bin/bash
# round-2 GPU session A: new parity tests, the eager-PyTorch bar, the XREG weight-gradient variant
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_gpu.txt 2>&1
free -g > gpurun_out/a_host.txt; nproc >> gpurun_out/a_host.txt
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_ddp.py > gpurun_out/a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/a_tests.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_ddp.py > gpurun_out/a_tests_all.log 2>&1; echo "tests rc=$?" >> gpurun_out/a_tests_all.log
DFCSA_WGRAD_XREG=1 timeout 300 python -m pytest tests/test_gpu_conv.py -k wgrad -q > gpurun_out/a_xreg_tests.log 2>&1; echo "xreg rc=$?" >> gpurun_out/a_xreg_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --detail gpurun_out/a_shapes.json > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
DFCSA_WGRAD_XREG=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --detail gpurun_out/a_shapes_xreg.json > gpurun_out/a_bench_xreg.json 2> gpurun_out/a_bench_xreg.err
timeout 600 python bench.py --impl reference --steps 2 > gpurun_out/a_ref.json 2> gpurun_out/a_ref.err
tail -3 gpurun_out/a_tests.log gpurun_out/a_xreg_tests.log
head -c 600 gpurun_out/a_bench.json
