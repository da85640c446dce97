"""LRH launch-shape sweep (profiling only): cluster size x CTAs/SM x threads x prefetch depth, 128x512x512 tiles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from regda_b200.utils.local_region_homog import Homogenizer

def run(lab, reg, bound, env, steps=30):
    for k in ("REGDA_LRH_CLUSTER", "REGDA_LRH_PER_SM", "REGDA_LRH_THREADS", "REGDA_LRH_PREFETCH"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    hom = Homogenizer(percent=0.5, class_num=6, ignore_label=-1, region_bound=bound, strict=False)
    try:
        for _ in range(3):
            out = hom(lab, reg)
    except Exception as e:
        return None, str(e)[:80]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        hom(lab, reg)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(steps))
    return ts[len(ts) // 2], out

def main():
  for R in (500, 50, 2000):
      lab, reg = bench.lrh_inputs(torch.device("cuda"), R, 2333)
      bound = int(reg.max()) + 1
      ref = None
      for cl, per, th in ((8, 2, 512), (16, 3, 256), (16, 2, 256), (16, 2, 512), (16, 4, 256), (8, 2, 256)):
          for pf in (0, 1, 2):
              ms, out = run(lab, reg, bound, dict(REGDA_LRH_CLUSTER=cl, REGDA_LRH_PER_SM=per, REGDA_LRH_THREADS=th, REGDA_LRH_PREFETCH=pf))
              if ms is None:
                  print(f"R={R} cluster={cl} per_sm={per} threads={th} prefetch={pf}: not launchable ({out})", flush=True)
                  break
              if ref is None:
                  ref = out
              ok = torch.equal(ref, out)
              print(f"R={R} cluster={cl} per_sm={per} threads={th} prefetch={pf}: {ms*1e3:7.1f} us  {24*lab.numel()/ms/1e6/6547.5:.3f} of HBM peak  same={ok}", flush=True)


if __name__ == '__main__':
    main()
