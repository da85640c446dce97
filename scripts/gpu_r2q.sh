#!/bin/bash
# round 2: epilogue rework (paired TMEM loads, LDS.128 + FADD2/FFMA2 statistics, CTA-level statistics flush): trace + conv tests
mkdir -p gpurun_out
REGDA_PDL=1 timeout 120 ./scripts/conv_trace > gpurun_out/r2q_conv_trace_pdl1.txt 2>&1
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_layers_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2q_tests.txt
tail -3 gpurun_out/r2q_tests.txt
