#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_teacher_gpu.py tests/test_tools_gpu.py -x -q 2>&1 | tail -30 > gpurun_out/c14_tests.txt
cat gpurun_out/c14_tests.txt
