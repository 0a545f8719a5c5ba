"""Where do the microseconds of ONE host-batch Evaluate go?  (BASELINE.json configs[0]: a+b over a 1M-row batch.)
  GDV_TRACE=1 python tools/host_latency.py [rows]        phase timings on stderr (csrc/gdv_runtime.cc)
Prints the median call latency for pageable (numpy) and pinned (gdv_host_alloc) host buffers."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import pyarrow as pa  # noqa: E402

import cases  # noqa: E402
import gandiva_b200 as gandiva  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    b = gandiva.TreeExprBuilder()
    schema, outs_t, _ = cases.case_arith("add", pa.int32())(b)
    p = gandiva.make_projector(schema, [b.make_expression(outs_t[0][0], pa.field("c", pa.int32()))], None)
    rng = np.random.default_rng(1)

    def pinned(nbytes):
        ptr = C.c_void_p()
        gandiva._check(gandiva.lib.gdv_host_alloc(nbytes, C.byref(ptr)))
        return ptr.value

    for kind in ("pageable", "pinned"):
        nb = (n + 63) // 64 * 8
        if kind == "pageable":
            bufs = [np.empty(n, np.int32), np.empty(n, np.int32), np.empty(nb, np.uint8), np.empty(nb, np.uint8),
                    np.empty(n, np.int32), np.empty(nb, np.uint8)]
            ptrs = [x.ctypes.data for x in bufs]
        else:
            ptrs = [pinned(n * 4), pinned(n * 4), pinned(nb), pinned(nb), pinned(n * 4), pinned(nb)]
            bufs = [np.ctypeslib.as_array(C.cast(p_, C.POINTER(C.c_uint8)), shape=(s,)) for p_, s in
                    zip(ptrs, (n * 4, n * 4, nb, nb, n * 4, nb))]
        for x in bufs[:2]:
            x.view(np.int32)[:] = rng.integers(-2**30, 2**30, n, dtype=np.int32)
        for x in bufs[2:4]:
            x.view(np.uint8)[:] = rng.integers(0, 256, nb, dtype=np.uint8)
        cols = (gandiva.gdv_column_t * 2)()
        cols[0].validity, cols[0].values = ptrs[2], ptrs[0]
        cols[1].validity, cols[1].values = ptrs[3], ptrs[1]
        cb = gandiva.gdv_batch_t(n, 2, gandiva.GDV_MEM_HOST, cols)
        oc = (gandiva.gdv_out_column_t * 1)()
        oc[0].values, oc[0].validity = ptrs[4], ptrs[5]

        def call():
            gandiva._check(gandiva.lib.gdv_projector_evaluate(p._h, C.byref(cb), None, oc, 1, None, 0))
        for _ in range(5):
            call()
        lat = []
        for _ in range(40):
            t0 = time.perf_counter()
            call()
            lat.append(time.perf_counter() - t0)
        lat.sort()
        print("%s host buffers, %d rows: median %.0f us, min %.0f us, p90 %.0f us  (H2D %.2f MB, D2H %.2f MB)" % (
            kind, n, lat[len(lat) // 2] * 1e6, lat[0] * 1e6, lat[int(len(lat) * 0.9)] * 1e6, n * 8.25 / 1e6, n * 4.125 / 1e6), flush=True)


if __name__ == "__main__":
    main()
