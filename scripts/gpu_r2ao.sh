#!/bin/bash
# round 2: final check after the inference BatchNorm change: full GPU tests, smoke, bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r2ao_tests.txt 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2ao_tests.txt | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --no-extras > gpurun_out/r2ao_bench.json 2> gpurun_out/r2ao_bench.err; cut -c1-200 gpurun_out/r2ao_bench.json
