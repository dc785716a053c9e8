"""Summarise an .ncu-rep (read here, no GPU): python tools/ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.sum", "smsp__cycles_active.avg",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = {k: i for i, k in enumerate(hdr)}
for r in rows[2:]:
    print("=== ", r[idx["Kernel Name"]][:60], "grid", r[idx.get("Grid Size", 0)], "block", r[idx.get("Block Size", 0)])
    for k in KEYS:
        if k in idx:
            print(f"  {k:75s} {r[idx[k]]:>16s} {units[idx[k]]}")
    for k in hdr:
        if "issue_stalled" in k and "per_issue_active" in k:
            v = float(r[idx[k]] or 0)
            if v > 0.15:
                print(f"  stall {k.split('issue_stalled_')[1].split('_per_issue')[0]:30s} {v:8.2f}")
