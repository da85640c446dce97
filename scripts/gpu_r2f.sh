#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lrh_gpu.py tests/test_conv_gpu.py tests/test_layers_gpu.py tests/test_step_gpu.py -m gpu -q --maxfail=20 > gpurun_out/r2f_tests.txt 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2f_tests.txt | head -30
for r in 50 500 1000 2000 5000; do
  timeout 300 python bench.py --workload lrh --regions $r --steps 20 --warmup 4 --no-cpu > gpurun_out/r2f_bench_lrh$r.json 2>> gpurun_out/r2f_bench_lrh.err
  python -c "import json;d=json.load(open('gpurun_out/r2f_bench_lrh$r.json'));print($r, d['value'], d['roofline']['frac'])"
done
REGDA_LRH_HASH=0 timeout 300 python bench.py --workload lrh --regions 2000 --steps 20 --warmup 4 --no-cpu > gpurun_out/r2f_bench_lrh2000_nohash.json 2>> gpurun_out/r2f_bench_lrh.err
python -c "import json;d=json.load(open('gpurun_out/r2f_bench_lrh2000_nohash.json'));print('nohash 2000', d['value'], d['roofline']['frac'])"
# racecheck / memcheck of the LRH cluster kernels (remote shared-memory reductions, cross-CTA pushes), small batch
cat > /tmp/lrh_san.py <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, '.')
from regda_b200 import synth
from regda_b200.utils.local_region_homog import Homogenizer
from oracle import cbind
for n_regions in (200, 3000):
    reg = synth.region_maps(2, 512, 512, n_regions, device="cuda", seed=5)
    lab = synth.lrh_labels(reg, 6, -1, seed=6)
    out = Homogenizer(percent=0.5, class_num=6, ignore_label=-1)(lab, reg)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), cbind.lrh(lab.cpu().numpy(), reg.cpu().numpy(), 6, -1, 0.5))
print("lrh sanitizer workload ok")
PY
timeout 900 compute-sanitizer --tool memcheck --kernel-regex kns=lrh python /tmp/lrh_san.py > gpurun_out/r2f_sanitizer_memcheck_lrh.txt 2>&1; tail -4 gpurun_out/r2f_sanitizer_memcheck_lrh.txt
timeout 900 compute-sanitizer --tool racecheck --kernel-regex kns=lrh python /tmp/lrh_san.py > gpurun_out/r2f_sanitizer_racecheck_lrh.txt 2>&1; tail -4 gpurun_out/r2f_sanitizer_racecheck_lrh.txt
timeout 600 python scripts/profile_step.py --engine auto --out gpurun_out/r2f_step_profile.txt > /dev/null 2>&1
timeout 600 python bench.py --steps 20 --warmup 4 --no-cpu --no-extras > gpurun_out/r2f_bench_step.json 2> gpurun_out/r2f_bench_step.err; cut -c1-200 gpurun_out/r2f_bench_step.json
