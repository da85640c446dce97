#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_conv_gpu.py tests/test_step_gpu.py -x -q 2>&1 | tail -25 > gpurun_out/c9_tests.txt
cat gpurun_out/c9_tests.txt
python bench.py --steps 20 --warmup 4 --no-cpu > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err; cut -c1-200 gpurun_out/c9_bench.json; tail -3 gpurun_out/c9_bench.err
REGDA_FUSE_BN_BWD=0 python bench.py --steps 20 --warmup 4 --no-cpu > gpurun_out/c9_bench_nofuse.json 2>> gpurun_out/c9_bench.err; cut -c1-200 gpurun_out/c9_bench_nofuse.json
