#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, ONE GPU).  Numbers printed by runs under ncu are never bench values.
set -x
mkdir -p gpurun_out
# 1. LRH kernel, full set, one launch after warm-up; and the launch list of the LRH microbench (LRH launches only)
ncu --set full --clock-control none --import-source on -k regex:lrh_cluster -s 3 -c 1 -f -o gpurun_out/lrh_r500_round1_final python bench.py --workload lrh --regions 500 --steps 3 --no-cpu > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lrh_ -c 12 --csv --log-file gpurun_out/launches_lrh_round1.csv python bench.py --workload lrh --regions 500 --steps 5 --no-cpu > /dev/null 2>&1
# 2. launch list of one training step (eager launches, no graph, so every kernel is a plain launch); ~1750 launches per step
REGDA_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 7000 -c 1800 --csv --log-file gpurun_out/launches_step_round1.csv python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
# 3. the dominant conv kernels, full set: PPM fuse conv fprop (persistent), its wgrad, a layer3 dgrad
ncu --set full --clock-control none --import-source on -k regex:conv_persistent_kernel -s 4 -c 1 -f -o gpurun_out/conv_head_fprop_round1 python scripts/bench_conv.py --n 16 --only head.fuse > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_persistent -s 4 -c 1 -f -o gpurun_out/conv_head_wgrad_round1 python scripts/bench_conv.py --n 16 --only head.fuse --wgrad > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_persistent_kernel -s 4 -c 1 -f -o gpurun_out/conv_l3conv2_dgrad_round1 python scripts/bench_conv.py --n 16 --only l3.conv2 --dgrad > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:bn_bwd_apply -s 40 -c 1 -f -o gpurun_out/bn_bwd_apply_round1 python scripts/profile_step.py --resnet resnet50 > /dev/null 2>&1
ls -la gpurun_out | tail -14
