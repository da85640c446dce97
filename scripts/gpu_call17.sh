#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_align_gpu.py -x -q 2>&1 | tail -30 > gpurun_out/c17_tests.txt
cat gpurun_out/c17_tests.txt
