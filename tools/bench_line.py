#!/usr/bin/env python
"""One-line digest of a bench.py JSON line. Usage: bench_line.py file.json"""
import json
import sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
r = d["roofline"]
print(d["config"]["workload"].split(":")[0], f"N={d['n_gpus']}", round(d["value"], 1), "GB/s", round(d["events_per_s"] / 1e6, 1), "Mev/s", round(d["ms_per_step"], 3), "ms",
      "pipeline_frac", round(r["pipeline"]["frac"], 3), "e2e", round(d.get("e2e", {}).get("value", 0), 1),
      "cpu1T", d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("events_per_s"),
      d["config"]["bytes_total"], d["config"]["msgs_total"], {k: round(v, 3) for k, v in r["kernels_ms"].items()})
