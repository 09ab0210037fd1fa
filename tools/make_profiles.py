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
            "  python bench.py --gpus 1 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extras      (c5, 10 GiB)\n"
            "Per-launch times under ncu are serialised and cold-cache: compare SHARES with bench.py's kernels_ms, not absolutes.\n\n")
    f.write(f"{'kernel':28s} {'n':>3s} {'avg us':>10s} {'share':>7s} {'dram rd MB':>11s} {'dram wr MB':>11s}\n")
    for k, a in agg.items():
        n = a["n"] or 1
        f.write(f"{k:28s} {int(n):3d} {a['gpu__time_duration.sum'] / n:10.1f} {100 * a['gpu__time_duration.sum'] / tot:6.1f}% "
                f"{a['dram__bytes_read.sum'] / n / 1e6:11.1f} {a['dram__bytes_write.sum'] / n / 1e6:11.1f}\n")
        traffic[f"c5/{k.replace('etl::', '')}"] = int((a["dram__bytes_read.sum"] + a["dram__bytes_write.sum"]) / n)
        traffic[f"c5/{k.replace('etl::', '')}/serialised_us"] = round(a["gpu__time_duration.sum"] / n, 1)
    f.write(f"\nsum of per-launch averages: {sum(a['gpu__time_duration.sum'] / (a['n'] or 1) for a in agg.values()) / 1e3:.3f} ms per decode\n")

# ---- launch lists of the other workloads (one decode each, tools/run_decode.py)
def summarise(src, dst, title, key):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ci = {c: i for i, c in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or not r[ci["Metric Value"]]:
            continue
        k = r[ci["Kernel Name"]].split("(")[0]
        m, u = r[ci["Metric Name"]], r[ci["Metric Unit"]]
        v = float(r[ci["Metric Value"]].replace(",", ""))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
        a = agg.setdefault(k, collections.defaultdict(float))
        a[m] += v * scale
        if m == "gpu__time_duration.sum":
            a["n"] += 1
    tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
    with open(dst, "w") as f:
        f.write(title + "\n\n")
        f.write(f"{'kernel':28s} {'n':>3s} {'avg us':>10s} {'share':>7s} {'dram rd MB':>11s} {'dram wr MB':>11s}\n")
        for k, a in agg.items():
            n = a["n"] or 1
            f.write(f"{k:28s} {int(n):3d} {a['gpu__time_duration.sum'] / n:10.1f} {100 * a['gpu__time_duration.sum'] / tot:6.1f}% "
                    f"{a['dram__bytes_read.sum'] / n / 1e6:11.1f} {a['dram__bytes_write.sum'] / n / 1e6:11.1f}\n")
            traffic[f"{key}/{k.replace('etl::', '')}"] = int((a["dram__bytes_read.sum"] + a["dram__bytes_write.sum"]) / n)
        f.write(f"\nsum of per-launch averages: {sum(a['gpu__time_duration.sum'] / (a['n'] or 1) for a in agg.values()) / 1e3:.3f} ms per decode\n")


for wname, scale in (("c2", "1.0"), ("c3", "0.3"), ("c4", "0.3")):
    src = os.path.join(G, f"launches_{wname}.csv")
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, f"{rnd}_launches_{wname}.csv"))
        summarise(src, os.path.join(P, f"{rnd}_launches_{wname}.txt"),
                  f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 28 -c 14\n"
                  f"  python tools/run_decode.py {wname} {scale} 3      (third decode; scale {scale} of the BASELINE size)", f"{wname}@{scale}")
json.dump({"_source": f"profiles/{rnd}_launches_*.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch; c5 at 10 GiB from the bench command)", **traffic},
          open(os.path.join(P, "traffic.json"), "w"), indent=1)

# ---- full captures
for tag, cmd in (("pipeline_c5", "c5 0.5"), ("rows_c3", "c3 0.3"), ("rows_c4", "c4 0.3")):
    rep = os.path.join(G, f"{rnd}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, "0"], capture_output=True, text=True).stdout
    lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "50", "smp"], capture_output=True, text=True).stdout
    with open(os.path.join(P, f"{rnd}_{tag}_ncu_full.txt"), "w") as f:
        f.write("ncu --set full --import-source on --clock-control none -k regex:'k_rows|k_heavy|...' (tools/final_measure.sh)\n"
                f"  python tools/run_decode.py {cmd} 3      (third decode)\n\n")
        f.write(out)
        f.write("\n--- hottest source lines over the captured kernels (samples)\n")
        f.write(lines)
for name in (f"{rnd}_bench_n1.json",):
    if os.path.exists(os.path.join(G, name)):
        txt = [l for l in open(os.path.join(G, name)) if l.startswith("{")]
        open(os.path.join(P, name), "w").write(txt[-1])
print("profiles written:", sorted(os.listdir(P)))
