#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pixel_gpu.py tests/test_step_gpu.py -x -q 2>&1 | tail -3
python bench.py --workload align --steps 20 --warmup 4 --no-cpu > gpurun_out/final_bench_align.json 2> gpurun_out/final_bench_align.err; cut -c1-220 gpurun_out/final_bench_align.json; tail -2 gpurun_out/final_bench_align.err
python bench.py --steps 20 --warmup 4 --no-cpu > gpurun_out/final_bench_step_b.json 2>> gpurun_out/final_bench_align.err; cut -c1-160 gpurun_out/final_bench_step_b.json
