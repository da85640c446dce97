#!/bin/bash
# final-state evidence for profiles/: kernel table of the step (torch.profiler, eager launches) and the ncu launch list of one step
mkdir -p gpurun_out
timeout 200 python scripts/profile_step.py --engine tcgen05 --out gpurun_out/final_step_profile.txt > /dev/null 2>&1; echo "profile rc=$?"
REGDA_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 4900 -c 1400 --csv --log-file gpurun_out/final_launches_step.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > /dev/null 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/final_launches_step.csv
