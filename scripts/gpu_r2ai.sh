#!/bin/bash
# round 2 final: full GPU tests, smoke, reference arm, default bench line (with LRH / library / CPU sub-records), kernel table
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r2ai_tests.txt 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2ai_tests.txt | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2ai_bench_reference.json 2> /dev/null; cut -c1-200 gpurun_out/r2ai_bench_reference.json
timeout 900 python bench.py > gpurun_out/r2ai_bench_default.json 2> gpurun_out/r2ai_bench_default.err; cut -c1-300 gpurun_out/r2ai_bench_default.json
timeout 600 python bench.py --config L --no-extras > gpurun_out/r2ai_bench_configL.json 2> gpurun_out/r2ai_bench_configL.err; cut -c1-300 gpurun_out/r2ai_bench_configL.json
timeout 600 python bench.py --workload align --no-extras > gpurun_out/r2ai_bench_align.json 2> gpurun_out/r2ai_bench_align.err; cut -c1-300 gpurun_out/r2ai_bench_align.json
timeout 600 python scripts/profile_step.py --engine auto --out gpurun_out/r2ai_step_profile.txt > /dev/null 2>&1
