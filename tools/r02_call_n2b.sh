#!/bin/bash
# Two GPUs: the SelectionVector push in waves — parity on 2 ranks (one run per rank, and 3 waves), then the
# weak-scaling bench with 4 waves / 1 wave / 8 waves, and (one GPU) the pageable e2e after the copy-pool change.
set -x
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k peer_selection_push > gpurun_out/r02n2b_pytest.log 2>&1; tail -3 gpurun_out/r02n2b_pytest.log
run() {  # name, extra bench flags
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-cpu $2 > gpurun_out/r02n2b_$1.json 2> gpurun_out/r02n2b_$1.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02n2b_$1.json").read().strip().splitlines()[-1])
    print("$1", "ms/step %.3f" % d["ms_per_step"], "rows/s %.4g" % d["value"], "kernel_ms %.3f" % d["roofline"]["kernel_ms"], d["config"]["gather_check"], [round(x, 2) for x in d["per_step_ms"]])
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/r02n2b_$1.err").read()[-1500:])
PY
}
run waves4 ""
run waves1 "--waves 1"
run waves8 "--waves 8"
timeout 300 python bench.py --no-configs --no-cpu --steps 5 --warmup 3 > gpurun_out/r02n2b_e2e_n1.json 2> gpurun_out/r02n2b_e2e_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02n2b_e2e_n1.json')); print('e2e pinned %.4g pageable %.4g' % (d['e2e']['value'], d['e2e']['pageable']['value']))"
