#!/bin/bash
# round-2 GPU session AD (8 GPUs): the final tree at N = 8 - c2 (weak) and c4 (strong)
N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/ad_c2_n8.json 2> gpurun_out/ad_c2_n8.err
timeout 400 $TR bench.py --gpus $N --config c4 --steps 6 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/ad_c4_n8.json 2> gpurun_out/ad_c4_n8.err
for f in gpurun_out/ad_c2_n8.json gpurun_out/ad_c4_n8.json; do head -c 260 $f; echo; done
tail -n 3 gpurun_out/ad_c2_n8.err gpurun_out/ad_c4_n8.err
