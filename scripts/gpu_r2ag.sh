#!/bin/bash
# round 2: proxy fence before the BatchNorm-input box is re-requested (generic reads -> async-proxy write), N = 2048 test shapes
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r2ag_tests.txt 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2ag_tests.txt | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --no-extras > gpurun_out/r2ag_bench.json 2> gpurun_out/r2ag_bench.err
cut -c1-300 gpurun_out/r2ag_bench.json
timeout 600 python scripts/profile_step.py --engine auto --out gpurun_out/r2ag_step_profile.txt > /dev/null 2>&1
