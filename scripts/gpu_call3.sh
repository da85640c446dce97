#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -x -q 2>&1 | tail -5 > gpurun_out/c3_tests.txt
cat gpurun_out/c3_tests.txt
python scripts/bench_conv.py --n 16 --graph 20 --stats > gpurun_out/c3_conv_fprop_tma.txt 2>&1
tail -n 1 gpurun_out/c3_conv_fprop_tma.txt
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err; cut -c1-200 gpurun_out/c3_bench.json
