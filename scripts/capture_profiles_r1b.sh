#!/bin/bash
# ncu evidence for profiles/ (round 1, session 5): run under gpurun, ONE GPU.  Numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out
# dominant conv kernels, full set: PPM fuse conv fprop with fused BN statistics (TMA-store epilogue), a layer3 dgrad+BN-reduce, head wgrad (TMA reduce epilogue)
ncu --set full --clock-control none --import-source on -k regex:conv_persistent_kernel -s 4 -c 1 -f -o gpurun_out/r1b_conv_head_fprop_stats python scripts/bench_conv.py --n 16 --only head.fuse --stats > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_persistent -s 4 -c 1 -f -o gpurun_out/r1b_conv_head_wgrad python scripts/bench_conv.py --n 16 --only head.fuse --wgrad > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_persistent_kernel -s 4 -c 1 -f -o gpurun_out/r1b_conv_l3conv2_fprop_stats python scripts/bench_conv.py --n 16 --only l3.conv2 --stats > /dev/null 2>&1
# launch list of one training step (eager launches so that every kernel is a plain launch)
REGDA_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 1600 --csv --log-file gpurun_out/r1b_launches_step.csv python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
# warm kernel table (CUPTI) + per-shape conv tables + the bench lines
python scripts/profile_step.py --engine auto --out gpurun_out/r1b_step_profile.txt > /dev/null 2>&1
python scripts/bench_conv.py --n 16 --graph 20 --stats > gpurun_out/r1b_conv_fprop.txt 2>&1
python scripts/bench_conv.py --n 16 --graph 20 --dgrad > gpurun_out/r1b_conv_dgrad.txt 2>&1
python scripts/bench_conv.py --n 16 --graph 20 --wgrad > gpurun_out/r1b_conv_wgrad.txt 2>&1
python bench.py --steps 20 --warmup 4 > gpurun_out/r1b_bench_step.json 2> gpurun_out/r1b_bench_step.err
python bench.py --workload lrh --regions 500 --steps 20 --warmup 4 > gpurun_out/r1b_bench_lrh500.json 2> gpurun_out/r1b_bench_lrh.err
ls -la gpurun_out | grep r1b
