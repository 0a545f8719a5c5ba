#!/bin/bash
# Round-2 eleventh single-GPU call (the last): the whole parity suite with the rope-consumer plans and the wider concats,
# smoke(), and the bench line.
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r02k_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02k_pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/r02k_bench_n1.json 2> gpurun_out/r02k_bench_n1.err; tail -c 200 gpurun_out/r02k_bench_n1.json; tail -2 gpurun_out/r02k_bench_n1.err
