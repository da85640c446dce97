#!/bin/bash
mkdir -p gpurun_out
for p in 0 1 2; do
  REGDA_PDL=$p timeout 600 python bench.py --steps 20 --warmup 4 --no-cpu --no-extras > gpurun_out/r2k_bench_pdl$p.json 2>> gpurun_out/r2k_bench.err; echo "PDL=$p $(python -c "import json;d=json.load(open('gpurun_out/r2k_bench_pdl$p.json'));print(d['value'], d['ms_per_step'])")"
done
REGDA_PDL=2 python scripts/bench_conv.py --n 16 --graph 20 --stats --only l3. 2>&1 | cut -c1-110
python scripts/bench_conv.py --n 16 --graph 20 --stats --only l3. 2>&1 | cut -c1-110
