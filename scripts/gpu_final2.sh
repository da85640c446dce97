#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 4 > gpurun_out/final_bench_2gpu.json 2> gpurun_out/final_bench_2gpu.err
cut -c1-160 gpurun_out/final_bench_2gpu.json
CUDA_VISIBLE_DEVICES=0 ncu --set full --clock-control none --import-source on -k regex:conv_persistent_kernel -s 4 -c 1 -f -o gpurun_out/final_conv_l3conv3_dgrad python scripts/bench_conv.py --n 16 --only l3.conv3 --dgrad > /dev/null 2>&1
ls -la gpurun_out | grep final_conv
