#!/bin/bash
# Round-end measurement on one B200: parity tests, smoke, bench (ours + reference arm), ncu launch list of the
# bench command, ncu --set full of the pipeline kernels.  Outputs → gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json | cut -c1-700
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_bench_c5.csv python bench.py --gpus 1 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -2 gpurun_out/bench_under_ncu.log | cut -c1-200
ETL_DEAD_SERIAL=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_cells|k_walk|k_utf8_dead|k_frames|k_index" -c 5 -s 10 -f -o gpurun_out/r01_pipeline_c5 python tools/run_decode.py c5 0.5 3 2>&1 | tail -1 | cut -c1-200
