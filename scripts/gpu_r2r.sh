#!/bin/bash
# round 2: epilogue fine trace + pooling rewrite / tap chain: full GPU tests + bench
mkdir -p gpurun_out
REGDA_PDL=1 timeout 120 ./scripts/conv_trace > gpurun_out/r2r_conv_trace_pdl1.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail 8 2>&1 | tail -40 > gpurun_out/r2r_tests.txt
tail -3 gpurun_out/r2r_tests.txt
timeout 600 python bench.py --no-extras > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
cut -c1-400 gpurun_out/r2r_bench.json
timeout 600 python scripts/profile_step.py --engine auto --out gpurun_out/r2r_step_profile.txt > /dev/null 2>&1
head -50 gpurun_out/r2r_step_profile.txt
