#!/bin/bash
# First gpurun call of the next round: everything built while no GPU minutes were left.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/r02_gpu_plan.sh'
# 1. the -m gpu suite on the real device (new: Arrow C device interface, new functions, key-scan
#    filter, concurrent Evaluate);  2. config-4 sweep: cooperative scan vs key-scan filter;
# 3. ncu of the key-scan kernel and of the nullable Q6 filter (no capture exists for either);
# 4. the default bench line.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
# memcheck + racecheck of the kernels that have never run on hardware (small cases only)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_z_optin_kernels.py -m gpu -x -q \
  -k "key_scan_filter_dense or two_pass or filter_walk or virtual_strings or number_to_text or digests or cast_string or regexp or case_trig or misc_casts" > gpurun_out/r02_memcheck.log 2>&1; tail -4 gpurun_out/r02_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_z_optin_kernels.py -m gpu -x -q \
  -k "key_scan_filter_dense or two_pass or filter_walk" > gpurun_out/r02_racecheck.log 2>&1; tail -4 gpurun_out/r02_racecheck.log
GDV_STR_COMBOS="512,2,0;128,1,16;256,1,16;512,1,16;1024,1,16" python tools/bench_configs.py str > gpurun_out/r02_str_sweep.log 2>&1
tail -8 gpurun_out/r02_str_sweep.log
GDV_STR_COMBOS="512,1,16" ncu --set full --clock-control none --import-source on -k regex:gdv_filter_expr -c 1 \
  -o gpurun_out/r02_keyscan python tools/bench_configs.py str 16000000 1 > gpurun_out/r02_keyscan_ncu.log 2>&1
GDV_Q6_COMBOS="1024,2,0;512,2,0;512,4,0;256,0,0,3;256,8,0,3;128,8,0,3;512,4,0,3;256,2,0,0,4;256,4,0,0,4;512,2,0,0,2;128,2,0,0,8" python tools/sweep_q6.py 1000000000 0 > gpurun_out/r02_q6_sweep.log 2>&1; tail -4 gpurun_out/r02_q6_sweep.log
# nullable variant after the branch-free validity loads (round 1: 0.837 at 1024 x 2)
GDV_Q6_COMBOS="1024,2,0;512,2,0;512,4,0;256,0,0,3;256,8,0,3;128,8,0,3;256,2,0,0,4;256,4,0,0,4" python tools/sweep_q6.py 1000000000 10 > gpurun_out/r02_q6_nulls_sweep.log 2>&1; tail -4 gpurun_out/r02_q6_nulls_sweep.log
GDV_Q6_COMBOS="1024,2,0" ncu --set full --clock-control none --import-source on -k regex:gdv_filter_expr -c 1 \
  -o gpurun_out/r02_q6_nulls python tools/sweep_q6.py 200000000 10 > gpurun_out/r02_q6_nulls_ncu.log 2>&1
python bench.py --filter-loader 3 --no-e2e --no-cpu > gpurun_out/r02_bench_n1_twopass.json 2> gpurun_out/r02_bench_n1_twopass.err; tail -c 600 gpurun_out/r02_bench_n1_twopass.json
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 1500 gpurun_out/r02_bench_n1.json
