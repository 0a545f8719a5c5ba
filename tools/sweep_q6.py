"""Tuning sweep for the fused Q6 filter kernel (rows_per_thread x block_threads) on device-
resident synthetic lineitem.  Writes gpurun_out/sweep_q6.json.  Usage:
  python tools/sweep_q6.py [rows] [null_permille]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import cases  # noqa: E402
import gandiva_b200 as gandiva  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000_000
    nullp = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    dev = torch.device("cuda")
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        st = stream.cuda_stream
        ship = torch.empty(n, dtype=torch.int32, device=dev)
        disc = torch.empty(n, dtype=torch.float64, device=dev)
        qty = torch.empty(n, dtype=torch.float64, device=dev)
        vl = [torch.empty((n + 31) // 32, dtype=torch.int32, device=dev) if nullp else None for _ in range(3)]
        for kind, t, v in ((0, ship, vl[0]), (1, disc, vl[1]), (2, qty, vl[2])):
            gandiva.generate_lineitem(0, kind, 42, 0, n, t.data_ptr(), v.data_ptr() if v is not None else 0, nullp, st)
        out = torch.empty(n, dtype=torch.int32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int64, device=dev)
        cols = [(v.data_ptr() if v is not None else 0, t.data_ptr(), 0, 0) for t, v in zip((ship, disc, qty), vl)]
        results = []
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        # (block_threads, rows_per_thread, string_scan flags, chunks per warp W); 0 = engine default
        combos = [(0, 0, 0, 0)] + [(bt, rpt, 0, w) for (bt, w) in ((256, 4), (256, 2), (512, 2), (128, 8), (1024, 1), (256, 8))
                                   for rpt in (2, 4, 8)]
        if os.environ.get("GDV_Q6_COMBOS"):
            combos = [tuple(int(x) for x in c.split(",")) for c in os.environ["GDV_Q6_COMBOS"].split(";")]
        for combo in combos:
            bt, rpt, flags = combo[:3]
            walk = combo[3] if len(combo) > 3 else 0
            if True:
                cfg = gandiva.Configuration(rows_per_thread=rpt, block_threads=bt, string_scan=flags, stages=walk)
                b = gandiva.TreeExprBuilder()
                f = gandiva.make_filter(cases.Q6_SCHEMA, b.make_condition(cases.q6_condition(b)), cfg)
                for _ in range(3):
                    f.evaluate_device(n, cols, out.data_ptr(), n, "UINT32", st, cnt.data_ptr())
                count = f.sync(st)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 10
                e0.record(stream)
                for _ in range(reps):
                    f.evaluate_device(n, cols, out.data_ptr(), n, "UINT32", st, cnt.data_ptr())
                e1.record(stream)
                stream.synchronize()
                f.sync(st)
                ms = e0.elapsed_time(e1) / reps
                bytes_ = n * (20.0 + (3 / 8.0 if nullp else 0.0)) + 4.0 * count
                gbs = bytes_ / (ms * 1e-3) / 1e9
                info = f.kernel_info
                r = {"block_threads": info["block_threads"], "rows_per_thread": info["rows_per_thread"], "flags": flags,
                     "walk": walk, "asked": list(combo), "ms": ms, "gbs": gbs, "frac": gbs / peak,
                     "regs": info["regs"], "rows_per_s": n / (ms * 1e-3), "count": count}
                results.append(r)
                print(json.dumps(r), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"rows": n, "null_permille": nullp, "results": results},
              open(os.path.join(ROOT, "gpurun_out", "sweep_q6_%d.json" % nullp), "w"), indent=1)


if __name__ == "__main__":
    main()
