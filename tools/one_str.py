"""Runs the config-4 string filter a few times on device-resident data (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import cases, gandiva_b200 as gandiva
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
rpt = int(sys.argv[2]) if len(sys.argv) > 2 else 0
bt = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hb = cases.comment_batch(n, seed=42); arr = hb.column(0)
dev = torch.device("cuda"); stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    st = stream.cuda_stream
    offs = torch.from_numpy(np.frombuffer(arr.buffers()[1], dtype=np.int32)[: n + 1].copy()).to(dev)
    data = torch.from_numpy(np.frombuffer(arr.buffers()[2], dtype=np.uint8).copy()).to(dev)
    vld = torch.from_numpy(np.frombuffer(arr.buffers()[0], dtype=np.uint8).copy()).to(dev)
    out = torch.empty(n, dtype=torch.int32, device=dev); cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    b = gandiva.TreeExprBuilder()
    f = gandiva.make_filter(cases.COMMENT_SCHEMA, b.make_condition(cases.comment_condition(b)), gandiva.Configuration(rows_per_thread=rpt, block_threads=bt))
    for _ in range(4):
        f.evaluate_device(n, [(vld.data_ptr(), offs.data_ptr(), data.data_ptr(), 0)], out.data_ptr(), n, "UINT32", st, cnt.data_ptr())
    print(f.sync(st), f.kernel_info)
