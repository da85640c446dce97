#!/bin/bash
# round 2: launch list of whole steps; graph-replay timelines at PDL level 1 / 2 / 2 with late BatchNorm trigger; tests after the ABI bump
mkdir -p gpurun_out
REGDA_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 4900 -c 1400 --csv --log-file gpurun_out/r2ad_launches_step.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > /dev/null 2>&1
wc -l gpurun_out/r2ad_launches_step.csv
REGDA_PDL=1 timeout 300 python scripts/timeline_step.py --out gpurun_out/r2ad_timeline_pdl1.json 2>&1 | tail -1
REGDA_PDL=2 timeout 300 python scripts/timeline_step.py --out gpurun_out/r2ad_timeline_pdl2.json 2>&1 | tail -1
timeout 1800 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r2ad_tests.txt 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2ad_tests.txt | head -20
timeout 600 python bench.py --no-extras > gpurun_out/r2ad_bench.json 2> gpurun_out/r2ad_bench.err
cut -c1-300 gpurun_out/r2ad_bench.json
