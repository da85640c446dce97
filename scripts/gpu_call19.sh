#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/c19_tests.txt
cat gpurun_out/c19_tests.txt
for v in 8 4 16 8 4; do
REGDA_BN_BLOCKS_PER_SM=$v python bench.py --steps 20 --warmup 4 --no-cpu > gpurun_out/c19_bench_bn$v.json 2>> gpurun_out/c19_bench.err; echo "BN_BLOCKS=$v $(cut -c50-70 gpurun_out/c19_bench_bn$v.json)"
done
