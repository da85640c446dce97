#!/bin/bash
mkdir -p gpurun_out
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 20 --warmup 4 > gpurun_out/r2m_bench_${n}gpu.json 2> gpurun_out/r2m_bench_${n}gpu.err; cut -c1-260 gpurun_out/r2m_bench_${n}gpu.json; tail -2 gpurun_out/r2m_bench_${n}gpu.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 3 --warmup 1 --impl reference > gpurun_out/r2m_bench_ref_8gpu.json 2>/dev/null; cut -c1-200 gpurun_out/r2m_bench_ref_8gpu.json
