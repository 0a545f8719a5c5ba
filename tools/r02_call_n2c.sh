#!/bin/bash
# Two GPUs, exactly the driver's commands (defaults: e2e and CPU arm included), both arms.
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --impl reference --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02n2c_ref.json 2> gpurun_out/r02n2c_ref.err; tail -c 400 gpurun_out/r02n2c_ref.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02n2c_bench.json 2> gpurun_out/r02n2c_bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02n2c_bench.json").read().strip().splitlines()[-1])
    print("N=2 ms/step %.3f" % d["ms_per_step"], "rows/s %.4g" % d["value"], d["config"]["gather_check"], d["config"]["host_numa"])
    e = d["e2e"]; print("e2e %.4g rows/s pageable %.4g" % (e["value"], e["pageable"]["value"]), "launches", d["gpu_launches"], d["clocks"])
except Exception as ex:
    print("failed", ex); print(open("gpurun_out/r02n2c_bench.err").read()[-2500:])
PY
