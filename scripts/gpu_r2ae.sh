#!/bin/bash
# round 2 experiment: the two domains on two streams (forward and, through autograd's stream tracking, backward)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_layers_gpu.py -m gpu -q --maxfail 8 2>&1 | tail -3
timeout 600 python bench.py --no-extras > gpurun_out/r2ae_bench_pair.json 2> gpurun_out/r2ae_bench_pair.err
cut -c1-200 gpurun_out/r2ae_bench_pair.json
REGDA_PAIR_FORWARD=0 timeout 600 python bench.py --no-extras > gpurun_out/r2ae_bench_seq.json 2> gpurun_out/r2ae_bench_seq.err
cut -c1-200 gpurun_out/r2ae_bench_seq.json; tail -2 gpurun_out/r2ae_bench_seq.err
REGDA_TWO_STREAMS=1 timeout 600 python bench.py --no-extras > gpurun_out/r2ae_bench_two.json 2> gpurun_out/r2ae_bench_two.err
cut -c1-200 gpurun_out/r2ae_bench_two.json; tail -2 gpurun_out/r2ae_bench_two.err
