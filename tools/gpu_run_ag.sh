#!/bin/bash
# round-2 GPU session AG (1 GPU): smoke() and the GPU tests on the final library
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/ag_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/ag_smoke.log
tail -n 5 gpurun_out/ag_smoke.log
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_ddp.py > gpurun_out/ag_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/ag_tests.log
tail -n 3 gpurun_out/ag_tests.log
