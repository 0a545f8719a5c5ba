#!/bin/bash
# Round-2 third single-GPU call: whole parity suite (new: device-side count chain, C++ API on HBM batches,
# raising arguments), string-filter sweep over chunks per warp, captures of the default Q1 (TMA) / add /
# string kernels, the launch list of the bench command, the bench line.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02c_pytest_gpu.log
python tools/bench_configs.py str > gpurun_out/r02c_str_sweep.log 2>&1; tail -12 gpurun_out/r02c_str_sweep.log | cut -c1-420
cp gpurun_out/bench_str.json gpurun_out/r02c_bench_str.json
GDV_STR_COMBOS="0,0,0,0" timeout 600 ncu --set full --clock-control none --import-source on -k regex:gdv_filter_expr -c 1 \
  -o gpurun_out/r02c_str_keydriven python tools/bench_configs.py str 50000000 1 > gpurun_out/r02c_str_ncu.log 2>&1
GDV_Q1_COMBOS="0,0,0,0" timeout 600 ncu --set full --clock-control none --import-source on -k regex:gdv_project_expr -c 1 \
  -o gpurun_out/r02c_q1_tma python tools/bench_configs.py q1 67108864 1 > gpurun_out/r02c_q1_ncu.log 2>&1; tail -2 gpurun_out/r02c_q1_ncu.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gdv_project_expr -c 1 \
  -o gpurun_out/r02c_add_i32 python tools/sweep_project.py > gpurun_out/r02c_add_ncu.log 2>&1; tail -2 gpurun_out/r02c_add_ncu.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02c_bench_under_ncu.log 2>&1
python bench.py > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err; tail -c 1200 gpurun_out/r02c_bench_n1.json; tail -5 gpurun_out/r02c_bench_n1.err
