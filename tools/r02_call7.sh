#!/bin/bash
# Round-2 seventh single-GPU call: pageable e2e against the copy-pool knobs, all on one box (runs on different
# boxes differed by 1.6x: the workers floated over both sockets).
set -x
mkdir -p gpurun_out
lscpu | grep -E "Model name|Socket|NUMA|Thread|Core" > gpurun_out/r02g_lscpu.txt; cat gpurun_out/r02g_lscpu.txt
nvidia-smi topo -m 2>/dev/null | head -12 >> gpurun_out/r02g_lscpu.txt
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 200 python bench.py --no-configs --no-cpu --steps 3 --warmup 3 --rows 268435456 > gpurun_out/r02g_$name.json 2> gpurun_out/r02g_$name.err
  python -c "
import json; d=json.load(open('gpurun_out/r02g_$name.json')); e=d['e2e']; print('$name', 'pinned %.4g rows/s %.1f GB/s | pageable %.4g rows/s %.1f GB/s' % (e['value'], e['h2d_gbs'], e['pageable']['value'], e['pageable']['h2d_gbs']))" | tee -a gpurun_out/r02g_sweep.txt
}
run default A=1
run nopin GDV_STAGE_PIN=0
run t16 GDV_STAGE_THREADS=16
run t16_nopin GDV_STAGE_THREADS=16 GDV_STAGE_PIN=0
run t4 GDV_STAGE_THREADS=4
run t12 GDV_STAGE_THREADS=12
run c256 GDV_STAGE_CHUNK_KB=256
run c512 GDV_STAGE_CHUNK_KB=512
run c2048 GDV_STAGE_CHUNK_KB=2048
run default_again A=1
for t in 8 16; do GDV_STAGE_THREADS=$t python tools/host_latency.py 2>&1 | grep pageable | sed "s/^/threads=$t /"; done | tee -a gpurun_out/r02g_sweep.txt
