#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_layers_gpu.py tests/test_step_gpu.py -x -q 2>&1 | tail -4 > gpurun_out/c6_tests.txt
cat gpurun_out/c6_tests.txt
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/c6_bench_pdl.json 2> gpurun_out/c6_bench.err; cut -c1-200 gpurun_out/c6_bench_pdl.json
REGDA_PDL=0 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/c6_bench_nopdl.json 2>> gpurun_out/c6_bench.err; cut -c1-200 gpurun_out/c6_bench_nopdl.json
python scripts/profile_step.py --engine auto --out gpurun_out/c6_step_profile.txt > /dev/null 2> gpurun_out/c6_profile.err
