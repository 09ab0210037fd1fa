#!/bin/bash
# Where the dead-segment UTF-8 pass runs (ETL_DEAD_MODE: 2 after the tuple pass [default], 0 side stream from the index pass,
# 1 side stream from the tuple pass, 3 inside k_rows).  Prints the per-kernel event times of the last decodes.
for m in 2 0 1 3; do
  for wl in "c5 0.5" "c2 1.0"; do
    echo "== ETL_DEAD_MODE=$m $wl"
    ETL_DEAD_MODE=$m timeout 120 python tools/run_decode.py $wl 6 2>&1 | tail -n 2
  done
done
