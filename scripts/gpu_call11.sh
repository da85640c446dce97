#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_layers_gpu.py -x -q -k "fused_into" 2>&1 | tail -25 > gpurun_out/c11_tests.txt
cat gpurun_out/c11_tests.txt
