#!/bin/bash
# round 2: misc kernel pass (maxpool index math, refine reciprocal math, CE column hoist, stem row-padded patches, cells op): tests + bench + profile
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail 12 2>&1 | tail -40 > gpurun_out/r2v_tests.txt
tail -3 gpurun_out/r2v_tests.txt
timeout 600 python bench.py --no-extras > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
cut -c1-300 gpurun_out/r2v_bench.json
timeout 600 python scripts/profile_step.py --engine auto --out gpurun_out/r2v_step_profile.txt > /dev/null 2>&1
REGDA_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ppm_pool_bwd|ppm_pool_fwd|ppm_cells|classifier_bwd" -s 4 -c 4 -f -o gpurun_out/r2v_misc python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > /dev/null 2>&1
ls -la gpurun_out/r2v_misc.ncu-rep
