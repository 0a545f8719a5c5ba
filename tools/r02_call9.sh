#!/bin/bash
# Round-2 ninth single-GPU call: host batches in slices on two streams — tests, then the per-call latency with and
# without slicing (pinned and pageable, 1M and 32M rows).
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_host_staging_gpu.py -m gpu -x -q > gpurun_out/r02i_pytest.log 2>&1; tail -3 gpurun_out/r02i_pytest.log
for sr in default 0 131072 262144 524288; do
  if [ $sr = default ]; then unset GDV_HOST_SLICE_ROWS; else export GDV_HOST_SLICE_ROWS=$sr; fi
  python tools/host_latency.py 2>&1 | sed "s/^/slice_rows=$sr /" | tee -a gpurun_out/r02i_latency.txt
done
for sr in default 0 2097152 8388608; do
  if [ $sr = default ]; then unset GDV_HOST_SLICE_ROWS; else export GDV_HOST_SLICE_ROWS=$sr; fi
  python tools/host_latency.py 32000000 2>&1 | sed "s/^/slice_rows=$sr /" | tee -a gpurun_out/r02i_latency.txt
done
unset GDV_HOST_SLICE_ROWS
GDV_TRACE=1 python tools/host_latency.py 2>&1 | grep "gdv trace" | sed -n '20,22p;70,72p' | tee -a gpurun_out/r02i_latency.txt
