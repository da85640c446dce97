#!/bin/bash
# round 2: tile order chosen per problem (strided unless the channel-block count does not divide the grid) + CTA-combined flush
mkdir -p gpurun_out
timeout 120 ./scripts/conv_trace > gpurun_out/r2aa_conv_trace.txt 2>&1
grep -E "^==" gpurun_out/r2aa_conv_trace.txt | cut -c1-110
timeout 600 python scripts/bench_conv.py --n 16 --graph 20 --stats > gpurun_out/r2aa_bench_conv_fprop_stats.txt 2>&1
timeout 600 python scripts/bench_conv.py --n 16 --graph 20 --dgrad > gpurun_out/r2aa_bench_conv_dgrad.txt 2>&1
tail -21 gpurun_out/r2aa_bench_conv_fprop_stats.txt | cut -c1-150
tail -1 gpurun_out/r2aa_bench_conv_dgrad.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail 12 2>&1 | tail -40 > gpurun_out/r2aa_tests.txt
tail -3 gpurun_out/r2aa_tests.txt
timeout 600 python bench.py --no-extras > gpurun_out/r2aa_bench.json 2> gpurun_out/r2aa_bench.err
cut -c1-300 gpurun_out/r2aa_bench.json
timeout 600 python scripts/profile_step.py --engine auto --out gpurun_out/r2aa_step_profile.txt > /dev/null 2>&1
