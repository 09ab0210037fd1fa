#!/bin/bash
# Round-end measurement on one B200: bench (ours, full extras), ncu launch lists per workload, ncu --set full of the
# tuple-pass kernels.  Outputs → gpurun_out/ (scratch); tools/make_profiles.py turns them into profiles/.
# Usage: tools/final_measure.sh [rNN] [quick]
set -u
R=${1:-r02}
mkdir -p gpurun_out
if [ "${2:-}" != "quick" ]; then
  timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err
  tail -c 1500 gpurun_out/${R}_bench_n1.json; tail -3 gpurun_out/${R}_bench_n1.err
fi
# launch list of the bench command (C5) and of one decode of the other workloads
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
  --log-file gpurun_out/launches_bench_c5.csv python bench.py --gpus 1 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu.log 2>&1
tail -1 gpurun_out/bench_under_ncu.log | cut -c1-160
for wl in "c2 1.0" "c3 0.3" "c4 0.3"; do
  set -- $wl
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 27 -c 12 --csv \
    --log-file gpurun_out/launches_$1.csv python tools/run_decode.py $1 $2 3 > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log | cut -c1-200
done
# full captures: the tuple pass on C5 / C3 / C4, the dead pass and the index passes on C5
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"k_rows|k_heavy|k_utf8_dead|k_long_cells|k_records|k_chase" -s 14 -c 6 -f \
  -o gpurun_out/${R}_pipeline_c5 python tools/run_decode.py c5 0.5 3 2>&1 | tail -1 | cut -c1-200
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"k_rows|k_heavy" -s 4 -c 2 -f \
  -o gpurun_out/${R}_rows_c3 python tools/run_decode.py c3 0.3 3 2>&1 | tail -1 | cut -c1-200
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"k_rows|k_heavy|k_records|k_chase" -s 10 -c 4 -f \
  -o gpurun_out/${R}_rows_c4 python tools/run_decode.py c4 0.3 3 2>&1 | tail -1 | cut -c1-200
ls -la gpurun_out/${R}_*.ncu-rep
