#!/bin/bash
# round 2: is the mid-run statistics flush slow?  harness with one / two statistics groups, per-shape product timings
mkdir -p gpurun_out
REGDA_TRACE_ONE_GROUP=1 timeout 120 ./scripts/conv_trace > gpurun_out/r2y_conv_trace_onegroup.txt 2>&1
timeout 120 ./scripts/conv_trace > gpurun_out/r2y_conv_trace_twogroups.txt 2>&1
grep -E "^==.*stats" gpurun_out/r2y_conv_trace_onegroup.txt | cut -c1-110
grep -E "^==.*stats" gpurun_out/r2y_conv_trace_twogroups.txt | cut -c1-110
timeout 600 python scripts/bench_conv.py --n 16 --graph 20 --stats > gpurun_out/r2y_bench_conv_fprop_stats.txt 2>&1
timeout 600 python scripts/bench_conv.py --n 16 --graph 20 --dgrad > gpurun_out/r2y_bench_conv_dgrad.txt 2>&1
tail -22 gpurun_out/r2y_bench_conv_fprop_stats.txt
