#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pixel_gpu.py tests/test_layers_gpu.py tests/test_step_gpu.py -x -q 2>&1 | tail -8 > gpurun_out/c5_tests.txt
cat gpurun_out/c5_tests.txt
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err; cut -c1-200 gpurun_out/c5_bench.json
