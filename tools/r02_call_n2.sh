#!/bin/bash
# Round-2 two-GPU call: the rewritten SelectionVector push (16-byte peer stores, 256-thread CTAs that fit
# next to the persistent filter) — parity on 2 ranks, then the weak-scaling bench in a few shapes.
set -x
mkdir -p gpurun_out
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k peer_selection_push > gpurun_out/r02n2_pytest.log 2>&1; tail -3 gpurun_out/r02n2_pytest.log
run() {  # name, extra bench flags
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-cpu $2 > gpurun_out/r02n2_$1.json 2> gpurun_out/r02n2_$1.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02n2_$1.json").read().strip().splitlines()[-1])
    print("$1", "ms/step %.3f" % d["ms_per_step"], "rows/s %.4g" % d["value"], "kernel_ms %.3f" % d["roofline"]["kernel_ms"], d["config"]["gather_check"], [round(x, 2) for x in d["per_step_ms"]])
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/r02n2_$1.err").read()[-1500:])
PY
}
run push_default ""
run push_c4_r1 "--push-ctas 4 --sm-reserve 1"
run push_c16_r3 "--push-ctas 16 --sm-reserve 3"
run push_c8_r0 "--push-ctas 8 --sm-reserve 0"
run nccl "--gather nccl"
run nogather "--no-gather"
run push_10b "--rows 1250000000"
