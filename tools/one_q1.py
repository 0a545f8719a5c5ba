"""Runs the Q1 projector a few times on device-resident data (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyarrow as pa, torch
import cases, gandiva_b200 as gandiva
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64 * 1024 * 1024
rpt = int(sys.argv[2]) if len(sys.argv) > 2 else 0
bt = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda"); stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    st = stream.cuda_stream; W = (n + 31) // 32; cols = []; keep = []
    for kind, f in zip(cases.Q1_KINDS, cases.Q1_SCHEMA):
        vals = torch.empty(n * (f.type.bit_width // 8), dtype=torch.uint8, device=dev)
        vld = torch.empty(W, dtype=torch.int32, device=dev)
        gandiva.generate_lineitem(0, kind, 42, 0, n, vals.data_ptr(), vld.data_ptr(), 20, st)
        keep += [vals, vld]; cols.append((vld.data_ptr(), vals.data_ptr(), 0, 0))
    outs = []
    for t in [pa.decimal128(32, 4), pa.decimal128(38, 6), pa.float64(), pa.float64(), pa.int64(), pa.float64(), pa.int64(), pa.int64()]:
        v = torch.empty(n * (t.bit_width // 8), dtype=torch.uint8, device=dev); vl = torch.empty(W, dtype=torch.int32, device=dev)
        keep += [v, vl]; outs.append((vl.data_ptr(), v.data_ptr()))
    b = gandiva.TreeExprBuilder()
    exprs = [b.make_expression(r, pa.field("o%d" % i, t)) for i, (r, t) in enumerate(cases.q1_outputs(b))]
    p = gandiva.make_projector(cases.Q1_SCHEMA, exprs, None, "NONE", gandiva.Configuration(rows_per_thread=rpt, block_threads=bt))
    for _ in range(4):
        p.evaluate_device(n, cols, outs, st)
    p.sync(st)
    print(p.kernel_info)
