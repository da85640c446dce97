#!/bin/bash
mkdir -p gpurun_out
for v in 2 3 4 2 3; do
REGDA_BN_BLOCKS_PER_SM=$v python bench.py --steps 20 --warmup 4 --no-cpu > gpurun_out/c20_bench_bn$v.json 2>> gpurun_out/c20_bench.err; echo "BN_BLOCKS=$v $(cut -c50-70 gpurun_out/c20_bench_bn$v.json)"
done
