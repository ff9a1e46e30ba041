#!/bin/bash
# round-2 GPU session AL (1 GPU): the FINAL tree - all single-GPU tests, smoke(), the default bench.py line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_ddp.py > gpurun_out/al_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/al_tests.log
tail -n 3 gpurun_out/al_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/al_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/al_smoke.log
tail -n 3 gpurun_out/al_smoke.log
timeout 400 python bench.py > gpurun_out/al_bench_default.json 2> gpurun_out/al_bench_default.err
head -c 300 gpurun_out/al_bench_default.json; echo; tail -n 2 gpurun_out/al_bench_default.err
