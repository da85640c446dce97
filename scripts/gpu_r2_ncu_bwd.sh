#!/bin/bash
# ncu --set full of the step's two largest backward kernels on layer3 shapes (profiles/ncu_conv_l3*_round2.txt)
mkdir -p gpurun_out
timeout 150 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_persistent -s 4 -c 1 -f -o gpurun_out/r2_l3conv2_wgrad python scripts/ncu_l3_bwd.py wgrad > gpurun_out/ncu_wgrad.log 2>&1; echo "wgrad rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:conv_persistent_kernel -s 4 -c 1 -f -o gpurun_out/r2_l3conv1_dgrad_bnred python scripts/ncu_l3_bwd.py bnred > gpurun_out/ncu_bnred.log 2>&1; echo "bnred rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
