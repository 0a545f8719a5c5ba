"""Host batches in PAGEABLE memory (what an Arrow RecordBatch from the reference's callers is) cross the
link through the engine's pinned staging ring (csrc/gdv_staging.cc); pinned host buffers and device
batches do not.  Same results either way, bit-exact against the oracle; `gdv_staged_bytes` is the
witness that the ring — not the driver's own pageable path — carried the bytes."""
import ctypes as C
import time

import numpy as np
import pyarrow as pa
import pytest

import cases
from helpers import assert_arrays_match

pytestmark = pytest.mark.gpu


def _add_projector(gandiva):
    b = gandiva.TreeExprBuilder()
    schema, outs, _ = cases.case_arith("add", pa.int32())(b)
    return b, schema, outs, gandiva.make_projector(schema, [b.make_expression(outs[0][0], pa.field("c", pa.int32()))],
                                                   None)


@pytest.mark.parametrize("n", [200_000, 1_000_000, 5_300_017])
def test_pageable_batches_go_through_the_ring(n, gandiva, oracle):
    b, schema, outs, p = _add_projector(gandiva)
    rng = np.random.default_rng(n)
    a = pa.array(rng.integers(-2**30, 2**30, n, dtype=np.int32), mask=rng.random(n) < 0.1)
    c = pa.array(rng.integers(-2**30, 2**30, n, dtype=np.int32), mask=rng.random(n) < 0.1)
    batch = pa.record_batch([a, c], schema=schema)
    before = gandiva.lib.gdv_staged_bytes()
    got = p.evaluate(batch)
    moved = gandiva.lib.gdv_staged_bytes() - before
    want = oracle.project([outs[0][0]], [pa.int32()], batch)
    assert_arrays_match(got[0], want[0])
    # copies of 1 MB and more are staged: two value columns in, one out (validity bitmaps are below the threshold
    # at the two smaller sizes)
    expect = 3 * 4 * n if 4 * n >= (1 << 20) else 0
    assert moved >= expect and (expect > 0 or moved == 0)


def test_back_to_back_and_after_a_pause(gandiva, oracle):
    """The copy threads go to sleep after 2 ms of idleness; calls right after each other, and calls after a
    pause (sleeping pool: the caller copies what no worker claims), give the same bytes."""
    b, schema, outs, p = _add_projector(gandiva)
    rng = np.random.default_rng(7)
    for it in range(24):
        n = int(rng.integers(300_000, 2_500_000))
        a = rng.integers(-2**30, 2**30, n, dtype=np.int32)
        c = rng.integers(-2**30, 2**30, n, dtype=np.int32)
        batch = pa.record_batch([pa.array(a), pa.array(c)], schema=schema)
        out = p.evaluate(batch)[0].to_numpy(zero_copy_only=False)
        assert np.array_equal(out, a + c), (it, n)
        if it % 3 == 0:
            time.sleep(float(rng.uniform(0.0, 0.008)))


def test_pinned_buffers_bypass_the_ring(gandiva):
    b, schema, outs, p = _add_projector(gandiva)
    n = 1_000_000
    nb = (n + 63) // 64 * 8
    ptrs = []
    for size in (4 * n, 4 * n, nb, nb, 4 * n, nb):
        q = C.c_void_p()
        gandiva._check(gandiva.lib.gdv_host_alloc(size, C.byref(q)))
        ptrs.append(q.value)
    try:
        view = lambda q, size, dt: np.ctypeslib.as_array(C.cast(q, C.POINTER(C.c_uint8)), shape=(size,)).view(dt)
        rng = np.random.default_rng(3)
        av, cv = view(ptrs[0], 4 * n, np.int32), view(ptrs[1], 4 * n, np.int32)
        av[:] = rng.integers(-2**30, 2**30, n, dtype=np.int32)
        cv[:] = rng.integers(-2**30, 2**30, n, dtype=np.int32)
        view(ptrs[2], nb, np.uint8)[:] = 0xFF
        view(ptrs[3], nb, np.uint8)[:] = 0xFF
        cols = (gandiva.gdv_column_t * 2)()
        cols[0].validity, cols[0].values = ptrs[2], ptrs[0]
        cols[1].validity, cols[1].values = ptrs[3], ptrs[1]
        cb = gandiva.gdv_batch_t(n, 2, gandiva.GDV_MEM_HOST, cols)
        oc = (gandiva.gdv_out_column_t * 1)()
        oc[0].values, oc[0].validity = ptrs[4], ptrs[5]
        before = gandiva.lib.gdv_staged_bytes()
        gandiva._check(gandiva.lib.gdv_projector_evaluate(p._h, C.byref(cb), None, oc, 1, None, 0))
        assert gandiva.lib.gdv_staged_bytes() == before
        assert np.array_equal(view(ptrs[4], 4 * n, np.int32), av + cv)
    finally:
        for q in ptrs:
            gandiva.lib.gdv_host_free(q)


def test_large_host_batch_of_every_buffer_kind(gandiva, oracle):
    """Host batches of 0.5M-1.3M rows through the staging ring: fixed-width, bool and string INPUTS, fixed-width
    and bool OUTPUTS, nulls, and columns that are themselves Arrow slices (offset != 0)."""
    I32, F64, B, S = pa.int32(), pa.float64(), pa.bool_(), pa.string()
    schema = pa.schema([("i", I32), ("j", I32), ("d", F64), ("p", B), ("s", S)])
    b = gandiva.TreeExprBuilder()
    f = {x.name: b.make_field(x) for x in schema}
    fn = b.make_function
    roots = [(fn("add", [f["i"], f["j"]], I32), I32),
             (fn("less_than", [f["d"], b.make_literal(0.25, F64)], B), B),
             (b.make_if(f["p"], fn("multiply", [f["d"], b.make_literal(2.0, F64)], F64), fn("castFLOAT8", [f["i"]], F64), F64),
              F64),
             (fn("octet_length", [f["s"]], I32), I32)]
    p = gandiva.make_projector(schema, [b.make_expression(r, pa.field("o%d" % k, t)) for k, (r, t) in enumerate(roots)], None)
    for n, off in ((524_288, 0), (700_001, 3), (1_300_000, 13)):
        batch = cases.random_batch(schema, n, seed=n, null_prob=0.15, offset=off)
        got = p.evaluate(batch)
        want = oracle.project([r for r, _ in roots], [t for _, t in roots], batch, threads=8)
        for k, (gv, wv) in enumerate(zip(got, want)):
            assert_arrays_match(gv, wv, "n=%d offset=%d output %d" % (n, off, k))


def test_error_late_in_a_large_host_batch_is_reported(gandiva):
    """divide by zero near the end of a 1.2M-row host batch raises, and the next call starts clean."""
    I32 = pa.int32()
    schema = pa.schema([("a", I32), ("b", I32)])
    b = gandiva.TreeExprBuilder()
    p = gandiva.make_projector(schema, [b.make_expression(
        b.make_function("divide", [b.make_field(schema.field(0)), b.make_field(schema.field(1))], I32), pa.field("q", I32))],
        None)
    n = 1_200_000
    a = np.arange(n, dtype=np.int32)
    d = np.full(n, 7, dtype=np.int32)
    ok = p.evaluate(pa.record_batch([pa.array(a), pa.array(d)], schema=schema))[0].to_numpy()
    assert np.array_equal(ok, a // 7)
    d[n - 5] = 0
    with pytest.raises(gandiva.GandivaError, match="divide by zero"):
        p.evaluate(pa.record_batch([pa.array(a), pa.array(d)], schema=schema))
    again = p.evaluate(pa.record_batch([pa.array(a), pa.array(np.full(n, 7, dtype=np.int32))], schema=schema))[0].to_numpy()
    assert np.array_equal(again, a // 7)        # the error flag was cleared
