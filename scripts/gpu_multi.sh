#!/bin/bash
# multi-GPU bench line of the self-training step: scripts/gpu_multi.sh N  (run under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 4 --no-extras > gpurun_out/multi_${N}gpu.json 2> gpurun_out/multi_${N}gpu.err
grep '^{' gpurun_out/multi_${N}gpu.json | tail -1 | cut -c1-400
