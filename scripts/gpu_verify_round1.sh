#!/bin/bash
# round-1 final: full GPU suite with the shipped defaults, the bench lines, the warm kernel table
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/final_tests.txt
cat gpurun_out/final_tests.txt
python bench.py --steps 20 --warmup 4 > gpurun_out/final_bench_step.json 2> gpurun_out/final_bench_step.err; cut -c1-160 gpurun_out/final_bench_step.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2>> gpurun_out/final_bench_step.err; cut -c1-200 gpurun_out/final_bench_reference.json
python scripts/profile_step.py --engine auto --out gpurun_out/final_step_profile.txt > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
