#!/usr/bin/env python
"""bench.py -- the measurement contract of this repo (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload step|lrh|align] [--impl ours|reference]

One JSON line on stdout (rank 0).  Under torchrun (N>1) every rank processes its own shard of
images (the path partitions by image, no data-path collective for LRH; gradient all-reduce for
the training step), timing is CUDA events bracketed by barrier + synchronize, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


# ----------------------------------------------------------------------------------------------
# plumbing
# ----------------------------------------------------------------------------------------------
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, pw = [], [], set(), []
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # the median of the busier half = clocks under load (idle samples sit at the low end)
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:]
        return {"sm_mhz": load[len(load) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
    return rank, world, local


def barrier(world):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(ms, world):
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return ms


# ----------------------------------------------------------------------------------------------
# workload: LRH microbench (BASELINE.json configs[4]): 128 x 512 x 512 tiles per GPU
# ----------------------------------------------------------------------------------------------
LRH_TILES, LRH_H, LRH_W = 128, 512, 512
LRH_BYTES_PER_PX = 24   # labels 8 + regions 8 + out 8


def lrh_inputs(device, n_regions, seed, tiles=LRH_TILES):
    from regda_b200 import synth
    reg = synth.region_maps(tiles, LRH_H, LRH_W, n_regions, device=device, seed=seed)
    lab = synth.lrh_labels(reg, 6, -1, seed=seed + 1)
    return lab, reg


def cpu_lrh_baseline(kind, tiles, n_regions, reps):
    """The CPU arm: the torch port of the reference's Homogenizer with every host thread
    (kind 'port'), on `tiles` tiles of the same synthetic workload."""
    from oracle import step_oracle as so
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    lab, reg = lrh_inputs("cpu", n_regions, 2333, tiles)
    so.lrh_torch(lab, reg, 6, -1, 0.5)
    t0 = time.perf_counter()
    for _ in range(reps):
        so.lrh_torch(lab, reg, 6, -1, 0.5)
    dt = (time.perf_counter() - t0) / reps
    return dict(value=tiles * LRH_H * LRH_W / dt / 1e9, unit="Gpix/s", cores=cores, kind=kind,
                sample=f"{tiles} tiles of 512x512, {n_regions} regions/tile, {reps} reps, torch port of Homogenizer.forward",
                ms_per_step=dt * 1e3)


def bench_lrh(args, rank, world, local):
    from regda_b200 import capi
    from regda_b200.utils.local_region_homog import Homogenizer
    dev = torch.device("cuda", local)
    pk = peaks()
    n_regions = args.regions
    lab, reg = lrh_inputs(dev, n_regions, 2333 + rank)
    bound = int(reg.max()) + 1
    if os.environ.get("REGDA_LRH_PATH"):
        capi.check(capi.lib().regda_set_lrh_path(int(os.environ["REGDA_LRH_PATH"])))
    hom = Homogenizer(percent=0.5, class_num=6, ignore_label=-1, region_bound=bound, strict=False)
    npx = lab.numel()

    # parity gate inside the bench: a sample of images against the C oracle
    out = hom(lab, reg)
    if rank == 0:
        import numpy as np
        from oracle import cbind
        for i in (0, LRH_TILES - 1):
            want = cbind.lrh(lab[i:i + 1].cpu().numpy(), reg[i:i + 1].cpu().numpy(), 6, -1, 0.5)
            assert np.array_equal(out[i:i + 1].cpu().numpy(), want), "LRH parity failure"
    hom.check()

    for _ in range(max(args.warmup, 3)):
        hom(lab, reg)
    sampler = ClockSampler(local)
    barrier(world)
    if rank == 0:
        sampler.start()
    launches0 = capi.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        hom(lab, reg)
        ev[i + 1].record()
    barrier(world)
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    launches = capi.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = max_over_ranks(total_ms, world)
    ms_step = total_ms / args.steps
    value = world * npx / (ms_step * 1e-3) / 1e9

    # e2e: host (pinned) int64 labels + regions -> device -> LRH -> host, every step
    h_lab, h_reg = lab.cpu().pin_memory(), reg.cpu().pin_memory()
    h_out = torch.empty_like(h_lab).pin_memory()
    d_lab, d_reg = torch.empty_like(lab), torch.empty_like(reg)

    def e2e_step():
        d_lab.copy_(h_lab, non_blocking=True)
        d_reg.copy_(h_reg, non_blocking=True)
        o = hom(d_lab, d_reg)
        h_out.copy_(o, non_blocking=True)

    for _ in range(2):
        e2e_step()
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(2, min(args.steps, 5))
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    barrier(world)
    e2e_ms = max_over_ranks(e0.elapsed_time(e1) / e2e_steps, world)
    assert torch.equal(h_out, out.cpu())

    if rank != 0:
        return
    ach = LRH_BYTES_PER_PX * npx / (sorted(per_launch)[len(per_launch) // 2] * 1e-3) / 1e9
    line = {
        "metric": "LRH throughput (Homogenizer.forward, int64 labels/regions)", "value": round(value, 3), "unit": "Gpix/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": f"LRH microbench: {LRH_TILES}x{LRH_H}x{LRH_W} tiles per GPU, ~{n_regions} regions/tile (region_bound {bound}), "
                               "6 classes, percent 0.5", "l2": "inputs (805 MB/GPU) larger than L2, no flush needed"},
        "e2e": {"value": round(world * npx / (e2e_ms * 1e-3) / 1e9, 3), "unit": "Gpix/s",
                "h2d_bytes_per_step": 16 * npx, "d2h_bytes_per_step": 8 * npx},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": round(ach, 1), "peak": pk["hbm"], "unit": "GB/s", "frac": round(ach / pk["hbm"], 4),
                     "traffic": None, "peak_source": pk["source"], "kernel": "lrh_cluster_kernel",
                     "algorithmic_bytes_per_launch": LRH_BYTES_PER_PX * npx},
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_lrh_baseline("port", 16, n_regions, 3)
    print(json.dumps(line), flush=True)


def reference_lrh(args, rank, world):
    if rank != 0:
        return
    steps, warm = args.steps, max(args.warmup, 1)
    tiles = int(os.environ.get("REGDA_REF_LRH_TILES", "16"))
    from oracle import step_oracle as so
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    lab, reg = lrh_inputs("cpu", args.regions, 2333, tiles)
    for _ in range(warm):
        so.lrh_torch(lab, reg, 6, -1, 0.5)
    t0 = time.perf_counter()
    for _ in range(steps):
        so.lrh_torch(lab, reg, 6, -1, 0.5)
    dt = (time.perf_counter() - t0) / steps
    v = tiles * LRH_H * LRH_W / dt / 1e9
    sample = f"each step = {tiles} of the {LRH_TILES} tiles (512x512, {args.regions} regions/tile), torch port of Homogenizer.forward"
    print(json.dumps({
        "impl": "reference", "metric": "LRH throughput (Homogenizer.forward, int64 labels/regions)", "value": round(v, 4),
        "unit": "Gpix/s", "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": round(dt * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": f"LRH microbench: {LRH_TILES}x{LRH_H}x{LRH_W} tiles per GPU, ~{args.regions} regions/tile"},
        "cpu_baseline": {"value": round(v, 4), "unit": "Gpix/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(v, 4), "unit": "Gpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=["step", "lrh", "align"])
    ap.add_argument("--regions", type=int, default=500, help="LRH microbench: regions per tile (50..5000)")
    ap.add_argument("--config", default="P", choices=["P", "L"], help="step workload: P = BASELINE.json configs[1] (the headline: ResNet-101, "
                    "512x512), L = configs[3] (ResNet-50, 7 classes, 1024x1024 tiles, 8 + 8 per GPU; quoted at 2 GPUs)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="step workload: skip the lrh sub-record and the GPU library baseline")
    args = ap.parse_args()
    rank, world, local = dist_setup(args.gpus)
    import bench_step as step_bench
    workload = args.workload or ("step" if step_bench is not None else "lrh")

    if args.impl == "reference":
        if workload == "lrh":
            reference_lrh(args, rank, world)
        else:
            step_bench.reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: regda_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        if workload == "lrh":
            bench_lrh(args, rank, world, local)
        else:
            step_bench.run(args, rank, world, local, peaks(), ClockSampler, barrier, max_over_ranks)
    finally:
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
