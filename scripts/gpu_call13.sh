#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_conv_gpu.py tests/test_layers_gpu.py tests/test_step_gpu.py tests/test_tools_gpu.py -x -q 2>&1 | tail -12 > gpurun_out/c13_tests.txt
cat gpurun_out/c13_tests.txt
python bench.py --steps 20 --warmup 4 --no-cpu > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err; cut -c1-200 gpurun_out/c13_bench.json; tail -3 gpurun_out/c13_bench.err
python scripts/profile_step.py --engine auto --out gpurun_out/c13_step_profile.txt > /dev/null 2> gpurun_out/c13_profile.err
