"""Launch loop of the bench line's dominant kernel (bench_step.head_conv_launcher) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_step
launch, flop = bench_step.head_conv_launcher()
for _ in range(8):
    launch()
torch.cuda.synchronize()
print("flop per launch", flop)
