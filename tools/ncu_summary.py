#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + hottest source lines) into text. Usage: ncu_summary.py rep [topN]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_global_ld.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
for r in rows[2:]:
    for h, u, v in zip(hdr, units, r):
        if h in want:
            print(f"{h} = {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
import collections, os
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, ""])
fname, ci = "?", None
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] in ("File Name", "File Path"):
        fname = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        ci = {c: i for i, c in enumerate(r) if c not in ("Source",)}
        ci["cuda"] = 1; ci["sass"] = 3
        continue
    if ci is None or len(r) < 8:
        continue
    try:
        k = (fname, int(r[0]))
    except ValueError:
        continue
    a = agg[k]
    try:
        a[0] += float(r[ci["# Samples"]]); a[1] += float(r[ci["Instructions Executed"]]); a[2] += float(r[ci["Thread Instructions Executed"]])
    except Exception:
        pass
    a[3] = r[1]
tot = sum(v[0] for v in agg.values()) or 1
toti = sum(v[1] for v in agg.values()) or 1
print(f"--- hottest source lines (samples total {tot:.0f}, warp-inst total {toti:.0f})")
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{100*v[0]/tot:5.1f}% smp {100*v[1]/toti:5.1f}% inst thr/inst {v[2]/max(v[1],1):5.1f}  {f}:{ln}: {v[3].strip()[:110]}")
