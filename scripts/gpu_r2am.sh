#!/bin/bash
# round 2: BatchNorm kernels with 16-byte loads of the per-channel constants (forward apply also requests its first batch first)
mkdir -p gpurun_out
timeout 120 ./scripts/bn_trace > gpurun_out/r2am_bn_trace.txt 2>&1; grep "^==" gpurun_out/r2am_bn_trace.txt
timeout 1800 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r2am_tests.txt 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2am_tests.txt | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2am_bench_default.json 2> gpurun_out/r2am_bench_default.err; cut -c1-300 gpurun_out/r2am_bench_default.json
timeout 600 python scripts/profile_step.py --engine auto --out gpurun_out/r2am_step_profile.txt > /dev/null 2>&1
