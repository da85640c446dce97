#!/bin/bash
# repeat the kernel-level parity tests (timing-dependent races show up as flaky failures)
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_layers_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -1; done | tee gpurun_out/soak.txt
