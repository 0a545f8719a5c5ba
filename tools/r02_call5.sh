#!/bin/bash
# Round-2 fifth single-GPU call: new tests (staging ring witness, push protocol with virtual ranks), host-call latency
# after the copy-pool change, bench line.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_selection_push_gpu.py tests/test_host_staging_gpu.py -m gpu -x -q > gpurun_out/r02e_pytest_new.log 2>&1; tail -3 gpurun_out/r02e_pytest_new.log
GDV_TRACE=1 python tools/host_latency.py > gpurun_out/r02e_host_latency.log 2>&1; grep -v "gdv trace" gpurun_out/r02e_host_latency.log; grep "gdv trace" gpurun_out/r02e_host_latency.log | sed -n 40,42p
for t in 1 2 4 16; do GDV_STAGE_THREADS=$t python tools/host_latency.py 2>&1 | grep pageable | sed "s/^/threads=$t /"; done | tee gpurun_out/r02e_host_latency_threads.log
python tools/host_latency.py 32000000 2>&1 | tee -a gpurun_out/r02e_host_latency_threads.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02e_pytest_gpu.log
python bench.py > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err; tail -c 600 gpurun_out/r02e_bench_n1.json; tail -3 gpurun_out/r02e_bench_n1.err
