"""Kernel timeline of CUDA-graph replays of the self-training step (torch.profiler / CUPTI): per kernel (name, stream, start, duration)
as compact JSON, to see where the replay's wall time goes -- kernel time per stream, idle gaps, overlap of the weight-gradient stream.
Usage: python scripts/timeline_step.py [--out gpurun_out/timeline.json]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="gpurun_out/timeline.json")
a = ap.parse_args()
import bench_step
model, step, runner, tensors = bench_step.build(torch.device("cuda", 0), 1, use_graph=True)
for _ in range(3):
    runner(*runner.static_in, lr=1e-2)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        runner(*runner.static_in, lr=1e-2)
    torch.cuda.synchronize()
tmp = a.out + ".chrome.json"
prof.export_chrome_trace(tmp)
ev = json.load(open(tmp))["traceEvents"]
ks = [dict(name=e["name"][:120], ts=e["ts"], dur=e["dur"], stream=e.get("args", {}).get("stream"), cat=e.get("cat"))
      for e in ev if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
ks.sort(key=lambda e: e["ts"])
json.dump(ks, open(a.out, "w"))
os.remove(tmp)
print(len(ks), "kernel events")
