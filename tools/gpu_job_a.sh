#!/bin/bash
# one combined measurement job (scratch → gpurun_out/)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "workload_parity or long_cell or hint or big_cells" 2>&1 | tail -3
ETL_DEAD_MODE=4 timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "workload_parity or long_cell or big_cells" 2>&1 | tail -3
bash tools/sweep_dead.sh 2>&1 | grep -v "^$" | cut -c1-330
timeout 200 ncu --set full --import-source on --clock-control none -k regex:"k_records|k_chase" -s 4 -c 2 -f -o gpurun_out/r02_records_c4 python tools/run_decode.py c4 0.3 3 2>&1 | tail -1 | cut -c1-200
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 24 --csv --log-file gpurun_out/launches_batch_c2.csv python tools/batch_probe.py c2 1.0 30 > /dev/null 2>&1
ETL_TRACE=1 timeout 100 python tools/batch_probe.py c2 1.0 300 2>&1 | tail -12
