"""Condense a torch.profiler key_averages table (scripts/profile_step.py output) to: kernel, total per step, avg, launches per step.
Usage: python scripts/summarize_profile.py gpurun_out/step_profile.txt [steps_profiled=2]"""
import re, sys
path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rows = []
for line in open(path):
    parts = re.split(r"\s{2,}", line.strip())
    if len(parts) < 8 or not parts[-1].isdigit():
        continue
    name = parts[0]
    if name.startswith(("aten::", "autograd::", "cuda", "cuLaunch", "Memcpy", "_", "torch::")) or "Backward" in name.split("(")[0] and "kernel" not in name:
        continue
    def us(t):
        v = float(re.sub(r"[a-z]+$", "", t)); return v * {"us": 1, "ms": 1e3, "s": 1e6}[re.search(r"[a-z]+$", t).group()]
    try:
        tot, avg = us(parts[-3]), us(parts[-2])
    except Exception:
        continue
    name = name.replace("regda::(anonymous namespace)::", "").replace("void ", "")
    name = re.sub(r"\(.*", "", name)[:70]
    rows.append((tot / steps, avg, int(parts[-1]) / steps, name))
rows.sort(reverse=True)
total = sum(r[0] for r in rows)
print(f"# per step: {total/1e3:.2f} ms of GPU kernel time in {sum(r[2] for r in rows):.0f} launches")
for tot, avg, n, name in rows[:int(sys.argv[3]) if len(sys.argv) > 3 else 45]:
    print(f"{tot:9.1f} us {100*tot/total:5.1f}%  x{n:5.0f}  avg {avg:8.1f} us  {name}")
