#!/bin/bash
# round 2: two-mode statistics flush; graph-replay timeline; ncu of the misc kernels; BatchNorm DRAM traffic in sequence
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_conv_gpu.py -m gpu -q --maxfail 8 2>&1 | tail -5 > gpurun_out/r2t_tests.txt
tail -2 gpurun_out/r2t_tests.txt
timeout 600 python bench.py --no-extras > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
cut -c1-300 gpurun_out/r2t_bench.json
timeout 600 python scripts/profile_step.py --engine auto --out gpurun_out/r2t_step_profile.txt > /dev/null 2>&1
timeout 600 python scripts/timeline_step.py --out gpurun_out/r2t_timeline.json 2>&1 | tail -2
REGDA_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ppm_pool_bwd|maxpool3s2|stem_im2col|ce_bilinear|refine_max|ppm_g_pack" -s 12 -c 10 -f -o gpurun_out/r2t_misc python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > /dev/null 2>&1
ls -la gpurun_out/r2t_misc.ncu-rep
REGDA_GRAPH=0 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --cache-control none --clock-control none -k regex:"bn_" -s 400 -c 400 --csv --log-file gpurun_out/r2t_bn_dram.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > /dev/null 2>&1
wc -l gpurun_out/r2t_bn_dram.csv
