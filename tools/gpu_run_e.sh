#!/bin/bash
# round-2 GPU session E (N GPUs, default 2): overlapped per-bucket all-reduce vs all buckets after the backward pass; C4
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
if [ "$N" == "2" ]; then
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/e_c2_n1.json 2> gpurun_out/e_c2_n1.err
fi
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/e_c2_n${N}_overlap.json 2> gpurun_out/e_c2_n${N}_overlap.err
DFCSA_DDP_OVERLAP=0 timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/e_c2_n${N}_after.json 2> gpurun_out/e_c2_n${N}_after.err
timeout 900 $TR bench.py --gpus $N --config c4 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/e_c4_n${N}.json 2> gpurun_out/e_c4_n${N}.err
for f in gpurun_out/e_c2_n1.json gpurun_out/e_c2_n${N}_overlap.json gpurun_out/e_c2_n${N}_after.json gpurun_out/e_c4_n${N}.json; do head -c 260 $f; echo; done
tail -n 3 gpurun_out/e_c4_n${N}.err
