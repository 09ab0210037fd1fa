#!/bin/bash
# Where the dead-segment UTF-8 pass runs (ETL_DEAD_MODE) x k_rows window (library variants built with ETL_LIB_SUFFIX /
# ETL_NVCC_DEFS, see etl_b200/build.py).  Prints the per-kernel event times of the last decode.
for cfg in ":2" ":4" "_b:4" "_b:2" ":0"; do
  v=${cfg%%:*}; m=${cfg##*:}
  for wl in "c5 0.5" "c2 1.0"; do
    echo "== lib '${v}' ETL_DEAD_MODE=$m $wl"
    ETL_LIB_SUFFIX=$v ETL_DEAD_MODE=$m timeout 120 python tools/run_decode.py $wl 6 2>&1 | tail -n 2
  done
done
