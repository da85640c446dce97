#!/bin/bash
# round 2: BatchNorm apply kernels with 16-byte loads of the per-channel statistics / affine parameters: phase trace, tests, bench
mkdir -p gpurun_out
timeout 120 ./scripts/bn_trace > gpurun_out/r2al_bn_trace_vec.txt 2>&1; cat gpurun_out/r2al_bn_trace_vec.txt
timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_step_gpu.py tests/test_bf16_parity_gpu.py -m gpu -q --maxfail 8 2>&1 | tail -3
timeout 600 python bench.py --no-extras > gpurun_out/r2al_bench.json 2> gpurun_out/r2al_bench.err; cut -c1-200 gpurun_out/r2al_bench.json
timeout 600 python scripts/profile_step.py --engine auto --out gpurun_out/r2al_step_profile.txt > /dev/null 2>&1
