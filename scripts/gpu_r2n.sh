#!/bin/bash
mkdir -p gpurun_out
for p in 0 -1 0 -1; do
  REGDA_MAIN_PRIORITY=$p timeout 600 python bench.py --steps 30 --warmup 4 --no-cpu --no-extras > gpurun_out/r2n_bench_prio$p.json 2>> gpurun_out/r2n_bench.err; echo "PRIO=$p $(python -c "import json;d=json.load(open('gpurun_out/r2n_bench_prio$p.json'));print(d['value'], d['ms_per_step'])")"
done
tail -3 gpurun_out/r2n_bench.err
