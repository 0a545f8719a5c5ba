"""Bandwidth of fused projector kernels on device-resident data (rows_per_thread x block_threads).
  python tools/sweep_project.py [rows]
Cases: add(int32,int32) without / with validity; Q6 predicate as a boolean projector."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import pyarrow as pa  # noqa: E402
import torch  # noqa: E402

import cases  # noqa: E402
import gandiva_b200 as gandiva  # noqa: E402


def timeit(fn, stream, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    stream.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    dev = torch.device("cuda")
    stream = torch.cuda.Stream()
    results = []
    with torch.cuda.stream(stream):
        st = stream.cuda_stream
        W = (n + 31) // 32
        a = torch.empty(n, dtype=torch.int32, device=dev)
        bcol = torch.empty(n, dtype=torch.int32, device=dev)
        av = torch.empty(W, dtype=torch.int32, device=dev)
        bv = torch.empty(W, dtype=torch.int32, device=dev)
        gandiva.generate_lineitem(0, 9, 42, 0, n, a.data_ptr(), av.data_ptr(), 100, st)
        gandiva.generate_lineitem(0, 10, 42, 0, n, bcol.data_ptr(), bv.data_ptr(), 100, st)
        out = torch.empty(n, dtype=torch.int32, device=dev)
        ov = torch.empty(W, dtype=torch.int32, device=dev)
        t = pa.int32()
        schema = pa.schema([("a", t), ("b", t)])
        # torch reference point: plain elementwise add (reads 8 B, writes 4 B per row)
        ms = timeit(lambda: torch.add(a, bcol, out=out), stream)
        results.append({"case": "torch.add int32 (reference point)", "ms": ms, "gbs": 12.0 * n / ms / 1e6})
        print(json.dumps(results[-1]), flush=True)
        for nulls in (False, True):
            for bt in (256, 512):
                for rpt in (2, 4, 8, 16):
                    b = gandiva.TreeExprBuilder()
                    root = b.make_function("add", [cases.F(b, "a", t), cases.F(b, "b", t)], t)
                    p = gandiva.make_projector(schema, [b.make_expression(root, pa.field("c", t))], None, "NONE",
                                               gandiva.Configuration(rows_per_thread=rpt, block_threads=bt))
                    cols = [(av.data_ptr() if nulls else 0, a.data_ptr(), 0, 0), (bv.data_ptr() if nulls else 0, bcol.data_ptr(), 0, 0)]
                    outs = [(ov.data_ptr(), out.data_ptr())]
                    ms = timeit(lambda: p.evaluate_device(n, cols, outs, st), stream)
                    p.sync(st)
                    bytes_ = n * (12.0 + (3 / 8.0 if nulls else 1 / 8.0))
                    r = {"case": "add_int32" + ("_nulls" if nulls else ""), "block_threads": bt, "rows_per_thread": rpt,
                         "ms": ms, "gbs": bytes_ / ms / 1e6, "frac": bytes_ / ms / 1e6 / peak, "regs": p.kernel_info["regs"]}
                    results.append(r)
                    print(json.dumps(r), flush=True)
        del a, bcol, av, bv, out, ov
        # Q6 predicate as a boolean projector: 20 B/row in, 2 bits/row out; no compaction
        m = min(n, 1_000_000_000)
        ship = torch.empty(m, dtype=torch.int32, device=dev)
        disc = torch.empty(m, dtype=torch.float64, device=dev)
        qty = torch.empty(m, dtype=torch.float64, device=dev)
        for kind, tn in ((0, ship), (1, disc), (2, qty)):
            gandiva.generate_lineitem(0, kind, 42, 0, m, tn.data_ptr(), 0, 0, st)
        ob = torch.empty((m + 31) // 32, dtype=torch.int32, device=dev)
        ovb = torch.empty((m + 31) // 32, dtype=torch.int32, device=dev)
        for bt in (256, 512):
            for rpt in (2, 4, 8):
                b = gandiva.TreeExprBuilder()
                p = gandiva.make_projector(cases.Q6_SCHEMA, [b.make_expression(cases.q6_condition(b), pa.field("k", pa.bool_()))],
                                           None, "NONE", gandiva.Configuration(rows_per_thread=rpt, block_threads=bt))
                cols = [(0, ship.data_ptr(), 0, 0), (0, disc.data_ptr(), 0, 0), (0, qty.data_ptr(), 0, 0)]
                ms = timeit(lambda: p.evaluate_device(m, cols, [(ovb.data_ptr(), ob.data_ptr())], st), stream)
                p.sync(st)
                bytes_ = m * 20.25
                r = {"case": "q6_predicate_bool_projector", "block_threads": bt, "rows_per_thread": rpt, "ms": ms,
                     "gbs": bytes_ / ms / 1e6, "frac": bytes_ / ms / 1e6 / peak, "regs": p.kernel_info["regs"]}
                results.append(r)
                print(json.dumps(r), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"rows": n, "results": results}, open(os.path.join(ROOT, "gpurun_out", "sweep_project.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
