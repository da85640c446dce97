#!/bin/bash
# round 2: 2-GPU sweep of NCCL's CTA budget (the all-reduce kernels share the SMs with the persistent convolutions)
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 4 --no-extras > gpurun_out/r2u_$tag.json 2> gpurun_out/r2u_$tag.err
  grep '^{' gpurun_out/r2u_$tag.json | tail -1 | cut -c1-220
}
run default X=1
run maxctas4 NCCL_MAX_CTAS=4
run maxctas8 NCCL_MAX_CTAS=8
run maxctas16 NCCL_MAX_CTAS=16
