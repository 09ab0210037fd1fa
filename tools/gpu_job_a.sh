#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for v in "" "_b"; do
  for wl in "c5 0.5" "c2 1.0" "c4 0.3"; do
    echo "== lib '${v}' $wl"
    ETL_LIB_SUFFIX=$v timeout 120 python tools/run_decode.py $wl 6 2>&1 | tail -n 1 | cut -c1-330
  done
done
timeout 100 python tools/batch_probe.py c2 1.0 300 notiming 2>&1 | tail -9 | head -6
