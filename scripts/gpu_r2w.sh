#!/bin/bash
# round 2: folded head with every weight part in place (no gather / scatter), g_pack v3, CE 3 blocks/SM, maxpool backward byte masks
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail 12 2>&1 | tail -40 > gpurun_out/r2w_tests.txt
tail -3 gpurun_out/r2w_tests.txt
timeout 600 python bench.py --no-extras > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
cut -c1-300 gpurun_out/r2w_bench.json
timeout 600 python scripts/profile_step.py --engine auto --out gpurun_out/r2w_step_profile.txt > /dev/null 2>&1
