#!/bin/bash
# round-2 (b): new conv paths (small maps, f32 parity, strided dgrad, classifier), bf16 parity, full suite, smoke, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -x --deselect tests/test_tools_gpu.py 2>&1 | tail -60 > gpurun_out/r2b_tests.txt
tail -40 gpurun_out/r2b_tests.txt
timeout 900 python -m pytest tests/test_tools_gpu.py -m gpu -q --maxfail=10 2>&1 | tail -30 > gpurun_out/r2b_tests_tools.txt
tail -15 gpurun_out/r2b_tests_tools.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 4 > gpurun_out/r2b_bench_step.json 2> gpurun_out/r2b_bench_step.err; cut -c1-300 gpurun_out/r2b_bench_step.json; tail -5 gpurun_out/r2b_bench_step.err
timeout 600 python scripts/profile_step.py --engine auto --out gpurun_out/r2b_step_profile.txt > /dev/null 2>&1
