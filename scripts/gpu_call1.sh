#!/bin/bash
# round-1 session-5 GPU call 1: warm per-kernel breakdown of the current step + per-shape conv timings without the host launch floor
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
python scripts/profile_step.py --engine auto --out gpurun_out/c1_step_profile.txt > /dev/null 2> gpurun_out/c1_profile.err
python scripts/bench_conv.py --n 16 --graph 20 --stats > gpurun_out/c1_conv_fprop.txt 2>&1
python scripts/bench_conv.py --n 16 --graph 20 --dgrad > gpurun_out/c1_conv_dgrad.txt 2>&1
python scripts/bench_conv.py --n 16 --graph 20 --wgrad > gpurun_out/c1_conv_wgrad.txt 2>&1
tail -3 gpurun_out/c1_conv_fprop.txt; cat gpurun_out/c1_bench.json
