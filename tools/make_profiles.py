#!/usr/bin/env python
"""Turns the scratch outputs of tools/final_measure.sh (gpurun_out/) into the tracked evidence under profiles/:
  profiles/rNN_launches_bench_c5.csv      ncu launch list of the bench command (raw)
  profiles/rNN_launches_bench_c5.txt      per-kernel averages / shares / DRAM bytes per launch
  profiles/rNN_pipeline_c5_ncu_full.txt   ncu --set full summaries + hottest source lines of the pipeline kernels
  profiles/traffic.json                   dram bytes (read+write) per launch, keyed "<workload>/<kernel>" (bench.py reads it)
Usage: python tools/make_profiles.py r01"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

# ---- launch list
src = os.path.join(G, "launches_bench_c5.csv")
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ci = {c: i for i, c in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[ci["Metric Value"]]:
        continue
    try:
        v = float(r[ci["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    k = r[ci["Kernel Name"]].split("(")[0]
    m, u = r[ci["Metric Name"]], r[ci["Metric Unit"]]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
    a = agg.setdefault(k, collections.defaultdict(float))
    a[m] += v * scale
    if m == "gpu__time_duration.sum":
        a["n"] += 1
shutil.copy(src, os.path.join(P, f"{rnd}_launches_bench_c5.csv"))
tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
traffic = {}
with open(os.path.join(P, f"{rnd}_launches_bench_c5.txt"), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none\n"
            "  python bench.py --gpus 1 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline      (c5, 10 GiB)\n"
            "Per-launch times under ncu are serialised and cold-cache: compare SHARES with bench.py's kernels_ms, not absolutes.\n\n")
    f.write(f"{'kernel':28s} {'n':>3s} {'avg us':>10s} {'share':>7s} {'dram rd MB':>11s} {'dram wr MB':>11s}\n")
    for k, a in agg.items():
        n = a["n"] or 1
        f.write(f"{k:28s} {int(n):3d} {a['gpu__time_duration.sum'] / n:10.1f} {100 * a['gpu__time_duration.sum'] / tot:6.1f}% "
                f"{a['dram__bytes_read.sum'] / n / 1e6:11.1f} {a['dram__bytes_write.sum'] / n / 1e6:11.1f}\n")
        traffic[f"c5/{k}"] = int((a["dram__bytes_read.sum"] + a["dram__bytes_write.sum"]) / n)
        traffic[f"c5/{k}/serialised_us"] = round(a["gpu__time_duration.sum"] / n, 1)
    f.write(f"\nsum of per-launch averages: {sum(a['gpu__time_duration.sum'] / (a['n'] or 1) for a in agg.values()) / 1e3:.3f} ms per decode\n")
json.dump({"_source": f"profiles/{rnd}_launches_bench_c5.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, c5 at 10 GiB)", **traffic},
          open(os.path.join(P, "traffic.json"), "w"), indent=1)

# ---- full captures
rep = os.path.join(G, f"{rnd}_pipeline_c5.ncu-rep")
if os.path.exists(rep):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, "0"], capture_output=True, text=True).stdout
    lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "40", "smp"], capture_output=True, text=True).stdout
    with open(os.path.join(P, f"{rnd}_pipeline_c5_ncu_full.txt"), "w") as f:
        f.write("ncu --set full --import-source on --clock-control none -k regex:'k_cells|k_walk|k_utf8_dead|k_frames|k_index'\n"
                "  ETL_DEAD_SERIAL=1 python tools/run_decode.py c5 0.5 3      (one decode of 5.4 GB, third iteration)\n\n")
        f.write(out)
        f.write("\n--- hottest source lines over the five kernels (samples)\n")
        f.write(lines)
print("profiles written:", sorted(os.listdir(P)))
