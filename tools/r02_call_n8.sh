#!/bin/bash
# Round-2 eight-GPU call: BASELINE.json configs[4] (10 B rows, 1.25e9 per GPU) exactly as the driver launches it,
# then N=4 on the same box.
set -x
mkdir -p gpurun_out
for N in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02n${N}_bench.json 2> gpurun_out/r02n${N}_bench.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02n${N}_bench.json").read().strip().splitlines()[-1])
    print("N=$N ms/step %.3f" % d["ms_per_step"], "rows/s %.4g" % d["value"], "kernel_ms %.3f" % d["roofline"]["kernel_ms"], d["config"]["rows_per_gpu"], d["config"]["gather_check"], [round(x, 2) for x in d["per_step_ms"]], "e2e", d["e2e"] and d["e2e"]["value"])
except Exception as e:
    print("N=$N failed", e); print(open("gpurun_out/r02n${N}_bench.err").read()[-2500:])
PY
done
