#!/bin/bash
# round 2: ncu --set full of BatchNorm apply launches inside the eager step (after the 16-byte constant loads)
mkdir -p gpurun_out
REGDA_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bn_apply_kernel|bn_bwd_apply_kernel" -s 300 -c 4 -f -o gpurun_out/r2aq_bn python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > /dev/null 2>&1
ls -la gpurun_out/r2aq_bn.ncu-rep
