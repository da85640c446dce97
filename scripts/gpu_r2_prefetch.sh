#!/bin/bash
# epilogue look-ahead prefetch: parity, kernel timing, one bench line
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_conv_gpu.py tests/test_layers_gpu.py -m gpu -x -q -p no:cacheprovider > gpurun_out/pf_tests.log 2>&1; echo "tests rc=$?"
tail -1 gpurun_out/pf_tests.log
timeout 100 python scripts/ncu_l3_bwd.py bnred time 2>&1 | tail -1
timeout 200 python bench.py --no-extras > gpurun_out/pf_bench.json 2> gpurun_out/pf_bench.err; cut -c1-330 gpurun_out/pf_bench.json
