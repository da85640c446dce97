#!/bin/bash
# round 2: head conv regression hunt: per-shape timings (A-then-B issue order restored)
mkdir -p gpurun_out
timeout 600 python scripts/bench_conv.py --n 16 --graph 20 --stats --only head > gpurun_out/r2ab_head.txt 2>&1
timeout 600 python scripts/bench_conv.py --n 16 --graph 20 --stats --only l4.conv > gpurun_out/r2ab_l4.txt 2>&1
timeout 600 python scripts/bench_conv.py --n 16 --graph 20 --only head > gpurun_out/r2ab_head_nostats.txt 2>&1
cat gpurun_out/r2ab_head.txt gpurun_out/r2ab_l4.txt gpurun_out/r2ab_head_nostats.txt | cut -c1-150
