#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/c12_tests.txt
cat gpurun_out/c12_tests.txt
python bench.py --steps 20 --warmup 4 --no-cpu > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err; cut -c1-200 gpurun_out/c12_bench.json; tail -3 gpurun_out/c12_bench.err
python scripts/profile_step.py --engine auto --out gpurun_out/c12_step_profile.txt > /dev/null 2> gpurun_out/c12_profile.err
