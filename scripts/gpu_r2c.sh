#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/r2c_tests.txt 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2c_tests.txt | head -60
timeout 300 python scripts/debug_bf16_parity.py resnet50 > gpurun_out/r2c_debug_r50.txt 2>&1; cat gpurun_out/r2c_debug_r50.txt | tail -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
