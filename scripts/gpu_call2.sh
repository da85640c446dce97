#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -x -q 2>&1 | tail -15 > gpurun_out/c2_tests.txt
cat gpurun_out/c2_tests.txt
python scripts/bench_conv.py --n 16 --graph 20 --stats > gpurun_out/c2_conv_fprop_tma.txt 2>&1
python scripts/bench_conv.py --n 16 --graph 20 --dgrad > gpurun_out/c2_conv_dgrad_tma.txt 2>&1
tail -1 gpurun_out/c2_conv_fprop_tma.txt gpurun_out/c2_conv_dgrad_tma.txt
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err; cut -c1-200 gpurun_out/c2_bench.json
