#!/bin/bash
mkdir -p gpurun_out
for cfg in "1 0" "1 1" "2 1"; do
  set -- $cfg
  REGDA_PDL=$1 REGDA_PDL_LATE=$2 timeout 600 python bench.py --steps 20 --warmup 4 --no-cpu --no-extras > gpurun_out/r2l_bench_pdl$1_late$2.json 2>> gpurun_out/r2l_bench.err; echo "PDL=$1 LATE=$2 $(python -c "import json;d=json.load(open('gpurun_out/r2l_bench_pdl$1_late$2.json'));print(d['value'], d['ms_per_step'])")"
done
timeout 1800 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r2l_tests.txt 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2l_tests.txt | head -20
