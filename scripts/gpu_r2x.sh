#!/bin/bash
# round 2: static-weight prefetch ahead of the PDL wait: phase trace with / without, tests, bench
mkdir -p gpurun_out
REGDA_TRACE_EARLY=0 timeout 120 ./scripts/conv_trace > gpurun_out/r2x_conv_trace_early0.txt 2>&1
REGDA_TRACE_EARLY=1 timeout 120 ./scripts/conv_trace > gpurun_out/r2x_conv_trace_early1.txt 2>&1
grep -E "^==" gpurun_out/r2x_conv_trace_early0.txt | cut -c1-110 | head -4
grep -E "^==" gpurun_out/r2x_conv_trace_early1.txt | cut -c1-110 | head -4
timeout 1500 python -m pytest tests -m gpu -q --maxfail 12 2>&1 | tail -40 > gpurun_out/r2x_tests.txt
tail -3 gpurun_out/r2x_tests.txt
timeout 600 python bench.py --no-extras > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
cut -c1-300 gpurun_out/r2x_bench.json
