#!/bin/bash
# Eight GPUs, 10 B rows: each step filtered in 4 waves (only the last wave's push is exposed).
set -x
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e --waves 4 > gpurun_out/r02n8c_bench.json 2> gpurun_out/r02n8c_bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02n8c_bench.json").read().strip().splitlines()[-1])
    print("N=8 waves4 ms/step %.3f" % d["ms_per_step"], "rows/s %.4g" % d["value"], "kernel_ms %.3f" % d["roofline"]["kernel_ms"], d["config"]["shard_rows"], d["config"]["gather_check"], [round(x, 2) for x in d["per_step_ms"]])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r02n8c_bench.err").read()[-2500:])
PY
