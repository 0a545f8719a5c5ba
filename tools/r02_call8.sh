#!/bin/bash
# Round-2 eighth single-GPU call: ring + copy workers on the GPU's NUMA node, feeding process bound there — pageable e2e
# five times on one box (variance), host-call latency, then the parity suite and the bench line with its reference arm.
set -x
mkdir -p gpurun_out
for i in 1 2 3 4 5; do
  timeout 200 python bench.py --no-configs --no-cpu --steps 3 --warmup 3 --rows 268435456 > gpurun_out/r02h_e2e_$i.json 2> gpurun_out/r02h_e2e_$i.err
  python -c "
import json; d=json.load(open('gpurun_out/r02h_e2e_$i.json')); e=d['e2e']; print('run $i', d['config']['host_numa'], '| pinned %.4g rows/s %.1f GB/s | pageable %.4g rows/s %.1f GB/s' % (e['value'], e['h2d_gbs'], e['pageable']['value'], e['pageable']['h2d_gbs']))" | tee -a gpurun_out/r02h_sweep.txt
done
timeout 200 python bench.py --no-configs --no-cpu --steps 3 --warmup 3 --rows 268435456 --no-numa-bind > gpurun_out/r02h_e2e_unbound.json 2> gpurun_out/r02h_e2e_unbound.err
python -c "
import json; d=json.load(open('gpurun_out/r02h_e2e_unbound.json')); e=d['e2e']; print('unbound process', '| pinned %.4g rows/s | pageable %.4g rows/s %.1f GB/s' % (e['value'], e['pageable']['value'], e['pageable']['h2d_gbs']))" | tee -a gpurun_out/r02h_sweep.txt
GDV_TRACE=1 python tools/host_latency.py 2>&1 | grep -v "gdv trace" | tee -a gpurun_out/r02h_sweep.txt
python tools/host_latency.py 32000000 2>&1 | tee -a gpurun_out/r02h_sweep.txt
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02h_pytest_gpu.log
timeout 420 python bench.py > gpurun_out/r02h_bench_n1.json 2> gpurun_out/r02h_bench_n1.err; tail -c 300 gpurun_out/r02h_bench_n1.json; tail -3 gpurun_out/r02h_bench_n1.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02h_bench_ref.json 2> gpurun_out/r02h_bench_ref.err; tail -c 300 gpurun_out/r02h_bench_ref.json
