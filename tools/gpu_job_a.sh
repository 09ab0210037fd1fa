#!/bin/bash
# one combined measurement job (scratch → gpurun_out/)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
ETL_DEAD_MODE=4 timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "workload_parity or long_cell or big_cells" 2>&1 | tail -2
for cfg in ":2" ":4" "_b:4"; do
  v=${cfg%%:*}; m=${cfg##*:}
  for wl in "c5 0.5" "c2 1.0" "c4 0.3"; do
    echo "== lib '${v}' ETL_DEAD_MODE=$m $wl"
    ETL_LIB_SUFFIX=$v ETL_DEAD_MODE=$m timeout 120 python tools/run_decode.py $wl 6 2>&1 | tail -n 1 | cut -c1-330
  done
done
timeout 100 python tools/batch_probe.py c2 1.0 300 2>&1 | tail -9
ETL_TRACE=1 timeout 100 python tools/batch_probe.py c2 1.0 300 notiming 2>&1 | tail -10
