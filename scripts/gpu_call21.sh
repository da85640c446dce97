#!/bin/bash
mkdir -p gpurun_out
run() { env "$@" python bench.py --steps 20 --warmup 4 --no-cpu 2>> gpurun_out/c21_bench.err | cut -c50-70; }
echo "base      $(run X=1)"
echo "unroll4   $(run REGDA_BN_UNROLL=4)"
echo "unroll4b2 $(run REGDA_BN_UNROLL=4 REGDA_BN_BLOCKS_PER_SM=2)"
echo "red2      $(run REGDA_BN_REDUCE_BLOCKS_PER_SM=2)"
echo "tiles149  $(run REGDA_CONV_MIN_TILES_256=149)"
echo "tiles37   $(run REGDA_CONV_MIN_TILES_256=37)"
echo "base      $(run X=1)"
timeout 600 python -m pytest tests/test_layers_gpu.py -x -q -k "bn_act or fused_into" 2>&1 | tail -2
REGDA_BN_UNROLL=4 timeout 600 python -m pytest tests/test_layers_gpu.py -x -q -k "bn_act or fused_into" 2>&1 | tail -2
