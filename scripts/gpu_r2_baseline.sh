#!/bin/bash
# round-2 start: GPU suite, bench line (label fix), LRH sweep over the config[4] range, warm kernel table
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2a_tests.txt
cat gpurun_out/r2a_tests.txt
timeout 600 python bench.py --steps 20 --warmup 4 > gpurun_out/r2a_bench_step.json 2> gpurun_out/r2a_bench_step.err; cut -c1-400 gpurun_out/r2a_bench_step.json
for r in 50 500 1000 2000 5000; do
  timeout 300 python bench.py --workload lrh --regions $r --steps 20 --warmup 4 --no-cpu > gpurun_out/r2a_bench_lrh$r.json 2>> gpurun_out/r2a_bench_lrh.err
  python -c "import json;d=json.load(open('gpurun_out/r2a_bench_lrh$r.json'));print($r, d['value'], d['roofline']['frac'])"
done
timeout 600 python scripts/profile_step.py --engine auto --out gpurun_out/r2a_step_profile.txt > /dev/null 2>&1
nvidia-smi --query-gpu=name,memory.total --format=csv
nproc
