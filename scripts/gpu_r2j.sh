#!/bin/bash
mkdir -p gpurun_out
for m in 0 129 300; do
  python scripts/bench_conv.py --n 16 --graph 20 --stats --min-tiles-256 $m > gpurun_out/r2j_conv_fprop_m$m.txt 2>&1
  python scripts/bench_conv.py --n 16 --graph 20 --dgrad --min-tiles-256 $m > gpurun_out/r2j_conv_dgrad_m$m.txt 2>&1
  echo "== min_tiles_256=$m"; grep -E "l3\.|l2\.conv|total" gpurun_out/r2j_conv_fprop_m$m.txt | cut -c1-110; grep -E "l3\.|total" gpurun_out/r2j_conv_dgrad_m$m.txt | cut -c1-110
done
python scripts/bench_conv.py --n 16 --graph 20 --wgrad > gpurun_out/r2j_conv_wgrad.txt 2>&1; grep total gpurun_out/r2j_conv_wgrad.txt
for m in 129 300; do
  REGDA_TUNE_MIN256=$m timeout 600 python bench.py --steps 20 --warmup 4 --no-cpu --no-extras > gpurun_out/r2j_bench_m$m.json 2>> gpurun_out/r2j_bench.err; cut -c1-160 gpurun_out/r2j_bench_m$m.json
done
