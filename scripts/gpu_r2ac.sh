#!/bin/bash
# round 2 verification: ncu of the roofline kernel, launch list, full tests, smoke, reference arm, default bench
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:conv_persistent_kernel -s 4 -c 1 -f -o gpurun_out/r2ac_conv_head_fold_fprop python scripts/ncu_head_conv.py > /dev/null 2>&1
ls -la gpurun_out/r2ac_conv_head_fold_fprop.ncu-rep
REGDA_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 6800 -c 1500 --csv --log-file gpurun_out/r2ac_launches_step.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > /dev/null 2>&1
wc -l gpurun_out/r2ac_launches_step.csv
timeout 1800 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r2ac_tests.txt 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2ac_tests.txt | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2ac_bench_reference.json 2> /dev/null; cut -c1-200 gpurun_out/r2ac_bench_reference.json
timeout 900 python bench.py > gpurun_out/r2ac_bench_default.json 2> gpurun_out/r2ac_bench_default.err; cut -c1-300 gpurun_out/r2ac_bench_default.json; tail -3 gpurun_out/r2ac_bench_default.err
timeout 600 python scripts/profile_step.py --engine auto --out gpurun_out/r2ac_step_profile.txt > /dev/null 2>&1
