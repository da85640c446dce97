#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r2i_tests.txt 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2i_tests.txt | head -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 4 --no-cpu --no-extras > gpurun_out/r2i_bench_step.json 2> gpurun_out/r2i_bench_step.err; cut -c1-200 gpurun_out/r2i_bench_step.json; tail -3 gpurun_out/r2i_bench_step.err
timeout 600 python scripts/profile_step.py --engine auto --out gpurun_out/r2i_step_profile.txt > /dev/null 2>&1
