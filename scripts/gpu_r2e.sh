#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/r2e_tests.txt 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2e_tests.txt | head -40
python scripts/debug_f32_accuracy.py 2>&1 | tail -6
timeout 900 python bench.py --steps 20 --warmup 4 --no-cpu --no-extras > gpurun_out/r2e_bench_step.json 2> gpurun_out/r2e_bench_step.err; cut -c1-200 gpurun_out/r2e_bench_step.json; tail -3 gpurun_out/r2e_bench_step.err
timeout 900 python bench.py --config L --steps 10 --warmup 3 --no-cpu > gpurun_out/r2e_bench_L.json 2> gpurun_out/r2e_bench_L.err; cut -c1-250 gpurun_out/r2e_bench_L.json; tail -3 gpurun_out/r2e_bench_L.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 20 --warmup 4 > gpurun_out/r2e_bench_2gpu.json 2> gpurun_out/r2e_bench_2gpu.err; cut -c1-200 gpurun_out/r2e_bench_2gpu.json; tail -3 gpurun_out/r2e_bench_2gpu.err
