#!/bin/bash
# round 2: epilogue v2 (single-segment TMEM loads, pack fast path), pooling backward v3, cells op, PDL level 2 with late BN trigger
mkdir -p gpurun_out
REGDA_PDL=1 timeout 120 ./scripts/conv_trace > gpurun_out/r2s_conv_trace_pdl1.txt 2>&1
timeout 1500 python -m pytest tests/test_layers_gpu.py tests/test_conv_gpu.py tests/test_step_gpu.py tests/test_bf16_parity_gpu.py -m gpu -q --maxfail 8 2>&1 | tail -40 > gpurun_out/r2s_tests.txt
tail -3 gpurun_out/r2s_tests.txt
timeout 600 python bench.py --no-extras > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
cut -c1-300 gpurun_out/r2s_bench.json
REGDA_PDL=2 REGDA_TUNE_BN_TRIGGER=0 timeout 600 python bench.py --no-extras > gpurun_out/r2s_bench_pdl2_late.json 2> gpurun_out/r2s_bench_pdl2_late.err
cut -c1-300 gpurun_out/r2s_bench_pdl2_late.json
REGDA_PDL=2 timeout 600 python bench.py --no-extras > gpurun_out/r2s_bench_pdl2.json 2> gpurun_out/r2s_bench_pdl2.err
cut -c1-300 gpurun_out/r2s_bench_pdl2.json
timeout 600 python scripts/profile_step.py --engine auto --out gpurun_out/r2s_step_profile.txt > /dev/null 2>&1
