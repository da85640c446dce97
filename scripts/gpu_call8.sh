#!/bin/bash
mkdir -p gpurun_out
python scripts/profile_step.py --engine auto --out gpurun_out/c8_step_profile.txt > /dev/null 2> gpurun_out/c8_profile.err
python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/c8_bench.json 2> gpurun_out/c8_bench.err; cut -c1-200 gpurun_out/c8_bench.json
