import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from regda_b200 import capi
from scripts.sweep_lrh import run
for R in (500, 50, 200, 1000):
    lab, reg = bench.lrh_inputs(torch.device("cuda"), R, 2333)
    bound = int(reg.max()) + 1
    ms, out = run(lab, reg, bound, {})
    print(f"R={R} auto (path {capi.lib().regda_lrh_last_path(None)}): {ms*1e3:7.1f} us  {24*lab.numel()/ms/1e6/6547.5:.3f}", flush=True)
    for cl, per, th in ((8, 2, 512), (16, 2, 512)):
        ms, out = run(lab, reg, bound, dict(REGDA_LRH_CLUSTER=cl, REGDA_LRH_PER_SM=per, REGDA_LRH_THREADS=th))
        if ms is not None:
            print(f"R={R} cluster={cl}: {ms*1e3:7.1f} us  {24*lab.numel()/ms/1e6/6547.5:.3f}", flush=True)
