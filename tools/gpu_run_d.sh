#!/bin/bash
# round-2 GPU session D (2 GPUs): NCCL correctness tests, what the collectives cost the step, C4 at N = 2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/d_gpus.txt
timeout 900 python -m pytest tests/test_gpu_ddp.py -m gpu -q > gpurun_out/d_ddp_tests.log 2>&1; echo "rc=$?" >> gpurun_out/d_ddp_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/d_c2_n1.json 2> gpurun_out/d_c2_n1.err
timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/d_c2_n2.json 2> gpurun_out/d_c2_n2.err
NCCL_MAX_CTAS=4 timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/d_c2_n2_cta4.json 2> gpurun_out/d_c2_n2_cta4.err
NCCL_MAX_CTAS=1 timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/d_c2_n2_cta1.json 2> gpurun_out/d_c2_n2_cta1.err
NCCL_MAX_CTAS=16 timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/d_c2_n2_cta16.json 2> gpurun_out/d_c2_n2_cta16.err
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING timeout 400 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/d_c2_n2_dbg.json 2> gpurun_out/d_c2_n2_dbg.err
timeout 900 $TR bench.py --gpus 2 --config c4 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/d_c4_n2.json 2> gpurun_out/d_c4_n2.err
tail -n 3 gpurun_out/d_ddp_tests.log
for f in gpurun_out/d_c2_n1.json gpurun_out/d_c2_n2.json gpurun_out/d_c2_n2_cta4.json gpurun_out/d_c2_n2_cta1.json gpurun_out/d_c2_n2_cta16.json gpurun_out/d_c4_n2.json; do head -c 260 $f; echo; done
