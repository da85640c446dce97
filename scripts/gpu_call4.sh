#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -x -q 2>&1 | tail -5 > gpurun_out/c4_tests.txt
cat gpurun_out/c4_tests.txt
python scripts/bench_conv.py --n 16 --graph 20 --wgrad > gpurun_out/c4_conv_wgrad_tma.txt 2>&1
tail -n 1 gpurun_out/c4_conv_wgrad_tma.txt
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err; cut -c1-200 gpurun_out/c4_bench.json
