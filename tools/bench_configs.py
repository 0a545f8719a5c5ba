"""Device-resident measurements of BASELINE.json configs[2] (Q1 eight-output projector) and
configs[3] (string filter like(upper(substr(c,1,32)), '%SPECIAL%REQUESTS%')).
  python tools/bench_configs.py q1 [rows_resident] [passes]
  python tools/bench_configs.py str [rows_per_batch] [batches]
Writes gpurun_out/bench_q1.json / bench_str.json.  CUDA-event timing, 3 warm-ups."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import pyarrow as pa  # noqa: E402
import torch  # noqa: E402

import cases  # noqa: E402
import gandiva_b200 as gandiva  # noqa: E402

PEAK = 6650.0
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]


def timeit(fn, stream, reps=5, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    stream.synchronize()
    return e0.elapsed_time(e1) / reps


def bench_q1(n, passes):
    dev = torch.device("cuda")
    stream = torch.cuda.Stream()
    results = []
    with torch.cuda.stream(stream):
        st = stream.cuda_stream
        W = (n + 31) // 32
        cols, keep = [], []
        for kind, f in zip(cases.Q1_KINDS, cases.Q1_SCHEMA):
            w = f.type.bit_width // 8
            vals = torch.empty(n * w, dtype=torch.uint8, device=dev)
            vld = torch.empty(W, dtype=torch.int32, device=dev)
            gandiva.generate_lineitem(0, kind, 42, 0, n, vals.data_ptr(), vld.data_ptr(), 20, st)
            keep += [vals, vld]
            cols.append((vld.data_ptr(), vals.data_ptr(), 0, 0))
        out_types = [pa.decimal128(32, 4), pa.decimal128(38, 6), pa.float64(), pa.float64(), pa.int64(),
                     pa.float64(), pa.int64(), pa.int64()]
        outs = []
        for t in out_types:
            v = torch.empty(n * (t.bit_width // 8), dtype=torch.uint8, device=dev)
            vl = torch.empty(W, dtype=torch.int32, device=dev)
            keep += [v, vl]
            outs.append((vl.data_ptr(), v.data_ptr()))
        in_bytes = 8 + 3 * 16 + 3 * 8 + 4 + 8 / 8.0
        out_bytes = 2 * 16 + 6 * 8 + 8 / 8.0
        combos = [(bt, rpt, 1, 0) for bt in (128, 256) for rpt in (2, 4)]
        combos += [(bt, rpt, 2, st_) for bt, rpt in ((128, 1), (128, 2), (256, 1), (256, 2), (128, 4), (512, 1))
                   for st_ in (2, 3, 4)]
        if os.environ.get("GDV_Q1_COMBOS"):
            combos = [tuple(int(x) for x in c.split(",")) for c in os.environ["GDV_Q1_COMBOS"].split(";")]
        for bt, rpt, loader, stages in combos:
            if True:
                b = gandiva.TreeExprBuilder()
                exprs = [b.make_expression(r, pa.field("o%d" % i, t)) for i, (r, t) in enumerate(cases.q1_outputs(b))]
                p = gandiva.make_projector(cases.Q1_SCHEMA, exprs, None, "NONE",
                                           gandiva.Configuration(rows_per_thread=rpt, block_threads=bt,
                                                                 loader=loader, stages=stages))

                def run():
                    for _ in range(passes):
                        p.evaluate_device(n, cols, outs, st)
                ms = timeit(run, stream)
                p.sync(st)
                rows = n * passes
                gbs = rows * (in_bytes + out_bytes) / ms / 1e6
                r = {"config": "q1_projector_8_outputs", "block_threads": bt, "rows_per_thread": rpt, "rows": rows,
                     "ms": ms, "rows_per_s": rows / ms * 1e3, "gbs": gbs, "frac": gbs / PEAK,
                     "bytes_per_row": in_bytes + out_bytes, "regs": p.kernel_info["regs"],
                     "loader": "tma" if p.kernel_info.get("staged") else "ldg",
                     "stages": p.kernel_info.get("stages"), "smem": p.kernel_info.get("dynamic_smem"),
                     "ctas_per_sm": p.kernel_info.get("blocks_per_sm")}
                results.append(r)
                print(json.dumps(r), flush=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "bench_q1.json"), "w"), indent=1)


def bench_str(rows_per_batch, batches):
    dev = torch.device("cuda")
    stream = torch.cuda.Stream()
    results = []
    block_rows = 2_000_000
    hb = cases.comment_batch(block_rows, seed=42)
    arr = hb.column(0)
    offs = np.frombuffer(arr.buffers()[1], dtype=np.int32)[: block_rows + 1].astype(np.int64)
    data = np.frombuffer(arr.buffers()[2], dtype=np.uint8)[: offs[-1]]
    vbits = np.frombuffer(arr.buffers()[0], dtype=np.uint8)[: block_rows // 8]
    reps = rows_per_batch // block_rows
    n = reps * block_rows
    with torch.cuda.stream(stream):
        st = stream.cuda_stream
        d_block = torch.from_numpy(data.copy()).to(dev)
        o_block = torch.from_numpy(offs.copy()).to(dev)
        v_block = torch.from_numpy(vbits.copy()).to(dev)
        block_bytes = int(offs[-1])
        assert block_bytes * reps < 2**31, "batch exceeds int32 offsets"
        d_bytes = d_block.repeat(reps)
        shifts = (torch.arange(reps, device=dev, dtype=torch.int64) * block_bytes).repeat_interleave(block_rows)
        d_offs = torch.empty(n + 1, dtype=torch.int32, device=dev)
        d_offs[:n] = (o_block[:block_rows].repeat(reps) + shifts).to(torch.int32)
        d_offs[n] = block_bytes * reps
        d_vld = v_block.repeat(reps)
        out = torch.empty(n, dtype=torch.int32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int64, device=dev)
        cols = [(d_vld.data_ptr(), d_offs.data_ptr(), d_bytes.data_ptr(), 0)]
        # (block_threads, rows_per_thread, string_scan[, stages]): string_scan 0 = key-driven filter (the default),
        # 4 = row-driven kernel; stages = 1024-row chunks per warp and tile of the key-driven kernel (0 = engine picks)
        combos = [(0, 0, 0, 0)] + [(bt, 0, 0, w) for bt in (128, 256, 512) for w in (1, 2, 4)] + [(512, 2, 4, 0)]
        if os.environ.get("GDV_STR_COMBOS"):
            combos = [tuple(int(x) for x in c.split(",")) for c in os.environ["GDV_STR_COMBOS"].split(";")]
        for combo in combos:
            bt, rpt, scan = combo[:3]
            stages = combo[3] if len(combo) > 3 else 0
            if True:
                b = gandiva.TreeExprBuilder()
                f = gandiva.make_filter(cases.COMMENT_SCHEMA, b.make_condition(cases.comment_condition(b)),
                                        gandiva.Configuration(rows_per_thread=rpt, block_threads=bt, string_scan=scan, stages=stages))

                def run():
                    for _ in range(batches):
                        f.evaluate_device(n, cols, out.data_ptr(), n, "UINT32", st, cnt.data_ptr())
                ms = timeit(run, stream)
                count = f.sync(st)
                rows = n * batches
                bytes_ = batches * (4.0 * n + block_bytes * reps + n / 8.0 + 4.0 * count)
                gbs = bytes_ / ms / 1e6
                r = {"config": "string_filter_like_upper_substr", "block_threads": f.kernel_info["block_threads"], "rows_per_thread": rpt,
                     "stages": stages, "tile_rows": f.kernel_info.get("tile_rows"),
                     "matcher": ("row-driven: " + ("per-lane" if scan & 1 else "cooperative scan") + ", cp.async prefetch") if scan & 4
                     else "key-driven: aligned-word scan of the column bytes",
                     "rows": rows, "ms": ms, "rows_per_s": rows / ms * 1e3, "gbs": gbs, "frac": gbs / PEAK,
                     "bytes_per_row": bytes_ / rows, "selected_per_batch": count, "regs": f.kernel_info["regs"],
                     "smem": f.kernel_info.get("dynamic_smem"), "ctas_per_sm": f.kernel_info.get("blocks_per_sm")}
                results.append(r)
                print(json.dumps(r), flush=True)
        # parity of the device-resident string path on the first block against the oracle
        import oracle
        b = gandiva.TreeExprBuilder()
        want = oracle.filter_indices(cases.comment_condition(b), hb, threads=8)
        got = out[:count]
        got = got[got < block_rows].cpu().numpy().astype(np.uint64)
        print(json.dumps({"string_parity_first_block": bool(np.array_equal(got, want)), "n": len(want)}), flush=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "bench_str.json"), "w"), indent=1)


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    which = sys.argv[1]
    if which == "q1":
        bench_q1(int(sys.argv[2]) if len(sys.argv) > 2 else 256 * 1024 * 1024, int(sys.argv[3]) if len(sys.argv) > 3 else 4)
    else:
        bench_str(int(sys.argv[2]) if len(sys.argv) > 2 else 50_000_000, int(sys.argv[3]) if len(sys.argv) > 3 else 2)
