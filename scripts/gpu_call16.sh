#!/bin/bash
mkdir -p gpurun_out
for v in 1 0 1 0; do
REGDA_WGRAD_STREAM=$v python bench.py --steps 20 --warmup 4 --no-cpu > gpurun_out/c16_bench_ws$v.json 2>> gpurun_out/c16_bench.err; echo "WS=$v $(cut -c50-70 gpurun_out/c16_bench_ws$v.json)"
done
tail -5 gpurun_out/c16_bench.err
timeout 900 python -m pytest tests/test_step_gpu.py tests/test_conv_gpu.py -x -q 2>&1 | tail -4
