#!/bin/bash
# Round-2 tenth single-GPU call: parity suite with the new functions (initcap, to_date, date-part aliases) and the final
# staging code, smoke(), then the bench line and its reference arm.
set -x
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/r02j_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02j_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 420 python bench.py > gpurun_out/r02j_bench_n1.json 2> gpurun_out/r02j_bench_n1.err; tail -c 300 gpurun_out/r02j_bench_n1.json; tail -3 gpurun_out/r02j_bench_n1.err
