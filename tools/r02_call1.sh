#!/bin/bash
# Round-2 first GPU call: measure everything that was built after the last GPU minute of round 1.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/r02_gpu.txt; nproc >> gpurun_out/r02_gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_z_optin_kernels.py -m gpu -x -q \
  -k "key_scan_filter_dense or two_pass or filter_walk" > gpurun_out/r02_memcheck.log 2>&1; tail -4 gpurun_out/r02_memcheck.log
GDV_STR_COMBOS="512,2,0;128,1,16;256,1,16;512,1,16;1024,1,16" python tools/bench_configs.py str > gpurun_out/r02_str_sweep.log 2>&1
tail -8 gpurun_out/r02_str_sweep.log
GDV_STR_COMBOS="512,1,16" timeout 600 ncu --set full --clock-control none --import-source on -k regex:gdv_filter_expr -c 1 \
  -o gpurun_out/r02_keyscan python tools/bench_configs.py str 16000000 1 > gpurun_out/r02_keyscan_ncu.log 2>&1
GDV_Q6_COMBOS="1024,2,0;512,2,0;512,4,0;256,0,0,3;256,8,0,3;128,8,0,3;512,4,0,3;256,2,0,0,4;256,4,0,0,4;512,2,0,0,2;128,2,0,0,8" python tools/sweep_q6.py 1000000000 0 > gpurun_out/r02_q6_sweep.log 2>&1; tail -14 gpurun_out/r02_q6_sweep.log
GDV_Q6_COMBOS="1024,2,0;512,2,0;512,4,0;256,0,0,3;256,8,0,3;128,8,0,3;256,2,0,0,4;256,4,0,0,4" python tools/sweep_q6.py 1000000000 10 > gpurun_out/r02_q6_nulls_sweep.log 2>&1; tail -10 gpurun_out/r02_q6_nulls_sweep.log
GDV_Q6_COMBOS="1024,2,0" timeout 600 ncu --set full --clock-control none --import-source on -k regex:gdv_filter_expr -c 1 \
  -o gpurun_out/r02_q6_nulls python tools/sweep_q6.py 200000000 10 > gpurun_out/r02_q6_nulls_ncu.log 2>&1
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 1500 gpurun_out/r02_bench_n1.json
