#!/bin/bash
# end-of-round verification on HEAD: the whole GPU suite, smoke, the default bench line (all legs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?"
tail -2 gpurun_out/final_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/final_bench.json
