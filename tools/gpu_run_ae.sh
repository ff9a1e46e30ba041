#!/bin/bash
# round-2 GPU session AE (8 GPUs): what the gradient exchange costs on the final tree - N = 1 on the same box, overlap vs after, a CTA cap
N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
A="--steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu"
timeout 300 python bench.py $A > gpurun_out/ae_n1.json 2> gpurun_out/ae_n1.err
timeout 300 $TR bench.py --gpus $N $A > gpurun_out/ae_n8_overlap.json 2> gpurun_out/ae_n8_overlap.err
DFCSA_DDP_OVERLAP=0 timeout 300 $TR bench.py --gpus $N $A > gpurun_out/ae_n8_after.json 2> gpurun_out/ae_n8_after.err
NCCL_MAX_CTAS=8 timeout 300 $TR bench.py --gpus $N $A > gpurun_out/ae_n8_overlap_cta8.json 2> gpurun_out/ae_n8_overlap_cta8.err
for f in n1 n8_overlap n8_after n8_overlap_cta8; do head -c 230 gpurun_out/ae_$f.json; echo; done
