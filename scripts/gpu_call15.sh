#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_tools_gpu.py -x -q 2>&1 | tail -8 > gpurun_out/c15_tests.txt
cat gpurun_out/c15_tests.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 4 > gpurun_out/c15_bench_2gpu.json 2> gpurun_out/c15_bench_2gpu.err
cut -c1-200 gpurun_out/c15_bench_2gpu.json; tail -3 gpurun_out/c15_bench_2gpu.err
