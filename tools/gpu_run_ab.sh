#!/bin/bash
# round-2 GPU session AB (2 GPUs): the final tree under data parallelism - NCCL correctness tests, c2 at N = 1 / 2 back to back, c4 at N = 2
N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
timeout 600 python -m pytest tests/test_gpu_ddp.py -m gpu -q > gpurun_out/ab_ddp_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/ab_ddp_tests.log
tail -n 3 gpurun_out/ab_ddp_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/ab_c2_n1.json 2> gpurun_out/ab_c2_n1.err
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/ab_c2_n2.json 2> gpurun_out/ab_c2_n2.err
timeout 600 $TR bench.py --gpus $N --config c4 --steps 4 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/ab_c4_n2.json 2> gpurun_out/ab_c4_n2.err
timeout 300 $TR bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/ab_ref_n2.json 2> gpurun_out/ab_ref_n2.err
for f in gpurun_out/ab_c2_n1.json gpurun_out/ab_c2_n2.json gpurun_out/ab_c4_n2.json gpurun_out/ab_ref_n2.json; do head -c 260 $f; echo; done
tail -n 3 gpurun_out/ab_c4_n2.err gpurun_out/ab_c2_n2.err
