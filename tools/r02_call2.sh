#!/bin/bash
# Round-2 second GPU call: the rewritten filter kernels (key-driven string filter, W chunks per warp,
# hoisted validity) on hardware: parity suite, memcheck of the new kernel, sweeps, ncu captures, new bench line.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02b_pytest_gpu.log
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_filter_kernels_gpu.py -m gpu -x -q \
  -k "key_driven_filter_dense or hoisted or walk_variant" > gpurun_out/r02b_memcheck.log 2>&1; tail -4 gpurun_out/r02b_memcheck.log
timeout 420 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_filter_kernels_gpu.py -m gpu -x -q \
  -k "key_driven_filter_dense" > gpurun_out/r02b_racecheck.log 2>&1; tail -4 gpurun_out/r02b_racecheck.log
python tools/bench_configs.py str > gpurun_out/r02b_str_sweep.log 2>&1; tail -9 gpurun_out/r02b_str_sweep.log
GDV_STR_COMBOS="0,0,0" timeout 600 ncu --set full --clock-control none --import-source on -k regex:gdv_filter_expr -c 1 \
  -o gpurun_out/r02b_str_keydriven python tools/bench_configs.py str 16000000 1 > gpurun_out/r02b_str_ncu.log 2>&1
python tools/sweep_q6.py 1000000000 0 > gpurun_out/r02b_q6_sweep.log 2>&1; tail -20 gpurun_out/r02b_q6_sweep.log
cp gpurun_out/sweep_q6_0.json gpurun_out/r02b_sweep_q6_nonull.json
python tools/sweep_q6.py 1000000000 10 > gpurun_out/r02b_q6_nulls_sweep.log 2>&1; tail -20 gpurun_out/r02b_q6_nulls_sweep.log
cp gpurun_out/sweep_q6_10.json gpurun_out/r02b_sweep_q6_nulls.json
GDV_Q6_COMBOS="0,0,0,0" timeout 600 ncu --set full --clock-control none --import-source on -k regex:gdv_filter_expr -c 1 \
  -o gpurun_out/r02b_q6_nulls python tools/sweep_q6.py 1000000000 10 > gpurun_out/r02b_q6_nulls_ncu.log 2>&1
GDV_Q6_COMBOS="0,0,0,0" timeout 600 ncu --set full --clock-control none --import-source on -k regex:gdv_filter_expr -c 1 \
  -o gpurun_out/r02b_q6 python tools/sweep_q6.py 1000000000 0 > gpurun_out/r02b_q6_ncu.log 2>&1
python bench.py > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; tail -c 3000 gpurun_out/r02b_bench_n1.json; tail -5 gpurun_out/r02b_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r02b_bench_ref.json 2> gpurun_out/r02b_bench_ref.err; tail -c 1500 gpurun_out/r02b_bench_ref.json
