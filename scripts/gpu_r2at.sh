#!/bin/bash
# round 2: last verification of the committed state: full GPU tests, smoke, default bench, kernel table
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r2at_tests.txt 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2at_tests.txt | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2at_bench_default.json 2> gpurun_out/r2at_bench_default.err; cut -c1-300 gpurun_out/r2at_bench_default.json
timeout 600 python scripts/profile_step.py --engine auto --out gpurun_out/r2at_step_profile.txt > /dev/null 2>&1
