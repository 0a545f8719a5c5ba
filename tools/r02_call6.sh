#!/bin/bash
# Round-2 sixth single-GPU call: push protocol with virtual ranks (opt-in test, under a timeout), the parity suite, bench line.
set -x
mkdir -p gpurun_out
GDV_TEST_VIRTUAL_RANKS=1 timeout 150 python -m pytest tests/test_selection_push_gpu.py tests/test_host_staging_gpu.py -m gpu -x -q -p no:xdist > gpurun_out/r02f_pytest_new.log 2>&1; tail -5 gpurun_out/r02f_pytest_new.log
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02f_pytest_gpu.log
timeout 420 python bench.py > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err; tail -c 600 gpurun_out/r02f_bench_n1.json; tail -3 gpurun_out/r02f_bench_n1.err
