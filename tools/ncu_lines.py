#!/usr/bin/env python
"""Per-source-line aggregation of an .ncu-rep source page. Usage: ncu_lines.py rep [topN] [sort=inst|smp|local]"""
import collections, csv, io, os, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40; sort = sys.argv[3] if len(sys.argv) > 3 else "inst"
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0.0, ""])
fname, ci = "?", None
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] in ("File Name", "File Path"): fname = os.path.basename(r[1]); continue
    if r[0] == "Line No": ci = {c: i for i, c in enumerate(r)}; continue
    if ci is None or len(r) < 20: continue
    if not r[0].strip(): continue            # SASS-only rows repeat the per-line totals
    a = agg[(fname, r[0])]
    try:
        a[0] += float(r[ci["# Samples"]] or 0); a[1] += float(r[ci["Instructions Executed"]] or 0); a[2] += float(r[ci["Thread Instructions Executed"]] or 0)
        if "Local" in r[ci["Address Space"]]: a[3] += float(r[ci["Instructions Executed"]] or 0)
    except Exception: pass
    a[4] = r[1]
ts = sum(v[0] for v in agg.values()) or 1; ti = sum(v[1] for v in agg.values()) or 1; tl = sum(v[3] for v in agg.values())
print(f"samples {ts:.0f}  warp-inst {ti:.0f}  local warp-inst {tl:.0f} ({100*tl/ti:.1f}%)  avg thr/inst {sum(v[2] for v in agg.values())/ti:.1f}")
key = {"inst": 1, "smp": 0, "local": 3}[sort]
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][key])[:topn]:
    print(f"{100*v[0]/ts:5.1f}% smp {100*v[1]/ti:5.1f}% inst loc {100*v[3]/ti:4.1f}% thr/inst {v[2]/max(v[1],1):5.1f}  {f}:{ln}: {v[4].strip()[:100]}")
