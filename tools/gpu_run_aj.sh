#!/bin/bash
# round-2 GPU session AJ (1 GPU): tree with the streaming-kernel changes of sessions AH / AI + the preloading block_out_fwd:
# all single-GPU tests, smoke(), step A/B of the block_out_fwd variant, a 20-step bench line, ncu --set full of the new row kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_ddp.py > gpurun_out/aj_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/aj_tests.log
tail -n 3 gpurun_out/aj_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/aj_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/aj_smoke.log
tail -n 3 gpurun_out/aj_smoke.log
B="python bench.py --warmup 3 --no-cpu-baseline --no-reference-gpu"
timeout 300 $B --steps 10 > gpurun_out/aj_bench_default.json 2> gpurun_out/aj_bench_default.err
DFCSA_BOUT_FWD_PRE=0 timeout 300 $B --steps 10 > gpurun_out/aj_bench_boutfwdpre0.json 2> gpurun_out/aj_bench_boutfwdpre0.err
timeout 300 $B --steps 20 --detail gpurun_out/aj_shapes.json > gpurun_out/aj_bench_steps20.json 2> gpurun_out/aj_bench_steps20.err
for f in default boutfwdpre0 steps20; do head -c 200 gpurun_out/aj_bench_$f.json; echo; tail -n 2 gpurun_out/aj_bench_$f.err; done
NB="python tools/block_bench.py --steps 1 --warmup 1 --level 1"
timeout 300 ncu --set full --clock-control none -k regex:"bilerpT_rows1|cols_reduce_par|pool_rows" -c 5 -f -o gpurun_out/aj_ncu_rows_l1 $NB > gpurun_out/aj_ncu.log 2>&1
tail -n 2 gpurun_out/aj_ncu.log
