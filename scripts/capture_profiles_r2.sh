#!/bin/bash
# ncu evidence for profiles/ (round 2): run under gpurun, ONE GPU.  Numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out
# LRH: hashed cluster kernel at 5000 regions/tile (why is it slow?), the global-bin path at 5000, the cluster kernel at 500 / 2000
ncu --set full --clock-control none --import-source on -k regex:lrh_cluster_fast -s 3 -c 1 -f -o gpurun_out/r2_lrh_r5000_hashed python bench.py --workload lrh --regions 5000 --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
REGDA_LRH_HASH=0 ncu --set full --clock-control none --import-source on -k regex:lrh_hist_global -s 3 -c 1 -f -o gpurun_out/r2_lrh_r5000_hist_global python bench.py --workload lrh --regions 5000 --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
REGDA_LRH_HASH=0 ncu --set full --clock-control none --import-source on -k regex:lrh_apply_global -s 3 -c 1 -f -o gpurun_out/r2_lrh_r5000_apply_global python bench.py --workload lrh --regions 5000 --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:lrh_cluster_fast -s 3 -c 1 -f -o gpurun_out/r2_lrh_r2000 python bench.py --workload lrh --regions 2000 --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:lrh_cluster_fast -s 3 -c 1 -f -o gpurun_out/r2_lrh_r500 python bench.py --workload lrh --regions 500 --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
# conv: PPM fuse conv fprop + stats (the bench line's dominant kernel), layer3 1x1 fprop, BatchNorm apply kernels
ncu --set full --clock-control none --import-source on -k regex:conv_persistent_kernel -s 4 -c 1 -f -o gpurun_out/r2_conv_head_fprop_stats python scripts/bench_conv.py --n 16 --only head.fuse --stats > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_persistent_kernel -s 4 -c 1 -f -o gpurun_out/r2_conv_l3conv1_fprop_stats python scripts/bench_conv.py --n 16 --only l3.conv1 --stats > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
