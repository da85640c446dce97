"""Print a window of a kernel timeline written by scripts/timeline_step.py: start (relative), duration, gap to the previous kernel
on the same stream.  Usage: python scripts/timeline_window.py timeline.json [first_index] [count]"""
import json, sys
ks = json.load(open(sys.argv[1]))
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
count = int(sys.argv[3]) if len(sys.argv) > 3 else 60
idx = [i for i, k in enumerate(ks) if "stem_im2col" in k["name"]]
base = idx[0] if idx else 0
one = ks[base:idx[1]] if len(idx) > 1 else ks[base:]
t0 = one[0]["ts"]
print("period us:", (ks[idx[1]]["ts"] - t0) if len(idx) > 1 else None, "kernels:", len(one))
last_end = {}
for i, k in enumerate(one[first:first + count]):
    st = k["stream"]
    gap = k["ts"] - last_end.get(st, k["ts"])
    name = k["name"].replace("void regda::(anonymous namespace)::", "").replace("regda::(anonymous namespace)::", "")[:60]
    print(f"{first + i:4d} s{st:<4} t={k['ts'] - t0:9.1f} dur={k['dur']:7.1f} gap={gap:6.1f}  {name}")
    last_end[st] = k["ts"] + k["dur"]
