#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into the few numbers DESIGN.md / bench.py quote.
Usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [band]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
band = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, v = rows[0], rows[-1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__average_warp_latency_per_inst_issued.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.avg",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed_op_shared_ld.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
for i, name in enumerate(h):
    if name in want or ("issue_stalled" in name and "per_issue_active" in name):
        try:
            val = float(v[i].replace(",", ""))
        except ValueError:
            continue
        if "issue_stalled" in name and val < 0.05:
            continue
        print(f"{name:85s} {v[i]} {rows[1][i]}")
if band:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
    h = rows[hi]
    data = [r for r in rows[hi + 1:] if len(r) == len(h)]
    ie, ws, si = h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)"), h.index("Source")

    def f(x):
        try:
            return float(x)
        except ValueError:
            return 0.0
    tot = sum(f(r[ie]) for r in data)
    tw = sum(f(r[ws]) for r in data)
    print(f"total warp instructions {tot:.4g}, stall samples {tw:.0f}, SASS rows {len(data)}")
    for b in range(0, len(data), band):
        s = sum(f(r[ie]) for r in data[b:b + band])
        w = sum(f(r[ws]) for r in data[b:b + band])
        if s / tot > 0.004 or w / tw > 0.004:
            print(f"{b:5d} inst={s / tot * 100:5.1f}% stall={w / tw * 100:5.1f}%  {data[b][si][:70]}")
