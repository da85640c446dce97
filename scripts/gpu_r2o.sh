#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:conv_persistent_kernel -s 4 -c 1 -f -o gpurun_out/r2_conv_head_fold_fprop python scripts/ncu_head_conv.py > /dev/null 2>&1
ls -la gpurun_out/r2_conv_head_fold_fprop.ncu-rep
# launch list of one training step (eager launches so that every kernel is a plain launch): kernel SHARES for profiles/
REGDA_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 7000 -c 1500 --csv --log-file gpurun_out/r2_launches_step.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > /dev/null 2>&1
wc -l gpurun_out/r2_launches_step.csv
timeout 1800 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r2o_tests.txt 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2o_tests.txt | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2o_bench_reference.json 2> /dev/null; cut -c1-200 gpurun_out/r2o_bench_reference.json
timeout 900 python bench.py > gpurun_out/r2o_bench_default.json 2> gpurun_out/r2o_bench_default.err; cut -c1-300 gpurun_out/r2o_bench_default.json; tail -3 gpurun_out/r2o_bench_default.err
