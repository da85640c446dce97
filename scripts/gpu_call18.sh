#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_align_gpu.py tests/test_tools_gpu.py -x -q 2>&1 | tail -14 > gpurun_out/c18_tests.txt
cat gpurun_out/c18_tests.txt
python tools/train_align_reg.py --config-path st.regda.2potsdam --steps 30 --sam-refine 2>&1 | tail -3 | tee gpurun_out/c18_align_2potsdam.txt
