#!/bin/bash
mkdir -p gpurun_out
for lvl in 0 1 2 0 1; do
REGDA_PDL=$lvl python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/c7_bench_pdl$lvl.json 2>> gpurun_out/c7_bench.err; echo "PDL=$lvl $(cut -c50-70 gpurun_out/c7_bench_pdl$lvl.json)"
done
