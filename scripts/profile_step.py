"""Kernel-time breakdown of one self-training step (torch.profiler / CUPTI; no nsys in the image).
Usage: python scripts/profile_step.py [--engine cudnn|auto] [--out gpurun_out/step_profile.txt]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity

ap = argparse.ArgumentParser()
ap.add_argument("--engine", default="cudnn")
ap.add_argument("--out", default="gpurun_out/step_profile.txt")
ap.add_argument("--resnet", default="resnet101")
a = ap.parse_args()
from regda_b200.ops import conv
conv.set_engine(a.engine)
import bench_step
model, step, runner, tensors = bench_step.build(torch.device("cuda", 0), 1, resnet=a.resnet, use_graph=False)
for _ in range(3):
    step(*tensors, 1e-2)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        step(*tensors, 1e-2)
    torch.cuda.synchronize()
tab = prof.key_averages().table(sort_by="cuda_time_total", row_limit=90, max_name_column_width=150)
os.makedirs(os.path.dirname(a.out), exist_ok=True)
open(a.out, "w").write(tab)
print(tab[-6000:])
