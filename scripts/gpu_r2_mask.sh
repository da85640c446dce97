#!/bin/bash
# mask-load vectorisation check: conv / layer parity, smoke, one bench line
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_conv_gpu.py tests/test_layers_gpu.py -m gpu -x -q > gpurun_out/mask_tests.log 2>&1; echo "tests rc=$?" 
tail -3 gpurun_out/mask_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 200 python bench.py --no-extras > gpurun_out/mask_bench.json 2> gpurun_out/mask_bench.err; cat gpurun_out/mask_bench.json
