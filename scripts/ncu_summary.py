"""Condense an `ncu --set full` capture to the lines profiles/ keeps.  Usage: python scripts/ncu_summary.py file.ncu-rep > profiles/x.txt"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__cluster_dim_x", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers"]
print(f"# ncu --set full --clock-control none summary of {rep.split('/')[-1]} (read with ncu -i ... --page raw --csv)")
for k in want:
    hit = [h for h in d if h == k or h.endswith("." + k)]
    for h in hit[:1]:
        print(f"{k} = {d[h][0]} {d[h][1]}")
for h in sorted(d):
    if "pipe_tensor" in h and "pct" in h and not any(h == k or h.endswith("." + k) for k in want):
        print(f"{h} = {d[h][0]} {d[h][1]}")
stall = {h.split("smsp__pcsamp_warps_issue_stalled_")[-1]: float(v[0].replace(",", "") or 0) for h, v in d.items() if "smsp__pcsamp_warps_issue_stalled_" in h and v[0] not in ("", "n/a")}
tot = sum(stall.values()) or 1
print("warp stall reasons (share of samples): " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(stall.items(), key=lambda kv: -kv[1])[:6]))
