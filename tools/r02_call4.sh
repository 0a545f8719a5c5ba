#!/bin/bash
# Round-2 fourth single-GPU call: parity suite with the final kernels, host-call latency breakdown, bench line.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02d_pytest_gpu.log
GDV_TRACE=1 python tools/host_latency.py > gpurun_out/r02d_host_latency.log 2>&1; grep -v "gdv trace" gpurun_out/r02d_host_latency.log; grep "gdv trace" gpurun_out/r02d_host_latency.log | tail -3; grep "gdv trace" gpurun_out/r02d_host_latency.log | sed -n 40,42p
for t in 1 4 16; do GDV_STAGE_THREADS=$t python tools/host_latency.py 2>&1 | grep pageable | sed "s/^/threads=$t /"; done | tee gpurun_out/r02d_host_latency_threads.log
python tools/host_latency.py 32000000 2>&1 | tee -a gpurun_out/r02d_host_latency_threads.log
for t in 12 16 24; do GDV_STAGE_THREADS=$t python tools/host_latency.py 32000000 2>&1 | grep pageable | sed "s/^/threads=$t /"; done | tee -a gpurun_out/r02d_host_latency_threads.log
python bench.py > gpurun_out/r02d_bench_n1.json 2> gpurun_out/r02d_bench_n1.err; tail -c 600 gpurun_out/r02d_bench_n1.json; tail -3 gpurun_out/r02d_bench_n1.err
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02d_bench_ref.json 2> gpurun_out/r02d_bench_ref.err; tail -c 400 gpurun_out/r02d_bench_ref.json
