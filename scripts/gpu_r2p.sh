#!/bin/bash
# round 2: phase trace of the persistent conv kernel on the layer1-4 shapes (scripts/conv_trace.cu)
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/r2p_clocks.txt
REGDA_PDL=1 timeout 120 ./scripts/conv_trace > gpurun_out/r2p_conv_trace_pdl1.txt 2>&1
REGDA_PDL=0 timeout 120 ./scripts/conv_trace > gpurun_out/r2p_conv_trace_pdl0.txt 2>&1
tail -5 gpurun_out/r2p_conv_trace_pdl1.txt
