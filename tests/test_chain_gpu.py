"""Filter -> Projector kept on the device (the reference's selection-vector
overload of Projector::Evaluate, P/includes/libgandiva.pxd:222-226, vector P/tests/test_gandiva.py:329-373):
a Q6-style Filter over lineitem writes its SelectionVector and its COUNT into device memory, the
Q1 eight-output Projector reads both from there on the same stream.  Nothing returns to the host
between the two Evaluate calls; the host learns the count at the very end.  Bit-exact against
oracle.filter_indices + oracle.project(selection=...)."""
import numpy as np
import pyarrow as pa
import pytest

import cases
import devmem
from helpers import assert_arrays_match

pytestmark = pytest.mark.gpu


def q6_on_q1_schema(b):
    """shipdate in [1994, 1995) AND 0.05 <= discount <= 0.07 AND quantity < 24 over the Q1 schema's columns."""
    f = {x.name: b.make_field(x) for x in cases.Q1_SCHEMA}
    B, SD, F64, I64 = pa.bool_(), pa.date32(), pa.float64(), pa.int64()
    fn = b.make_function
    return b.make_and([
        fn("greater_than_or_equal_to", [f["l_shipdate"], b.make_literal(8766, SD)], B),
        fn("less_than", [f["l_shipdate"], b.make_literal(9131, SD)], B),
        fn("greater_than_or_equal_to", [f["l_discount_f"], b.make_literal(0.05, F64)], B),
        fn("less_than_or_equal_to", [f["l_discount_f"], b.make_literal(0.07, F64)], B),
        fn("less_than", [f["l_quantity"], b.make_literal(24, I64)], B)])


@pytest.mark.parametrize("mode,npt", [("UINT32", np.uint32), ("UINT64", np.uint64)])
@pytest.mark.parametrize("n", [1, 4097, 300_011])
def test_filter_then_projector_without_host_round_trip(n, mode, npt, gandiva, oracle):
    b = gandiva.TreeExprBuilder()
    cond = q6_on_q1_schema(b)
    outs = cases.q1_outputs(b)
    batch = cases.q1_batch(n, seed=42, null_permille=20)
    filt = gandiva.make_filter(cases.Q1_SCHEMA, b.make_condition(cond))
    exprs = [b.make_expression(r, pa.field("o%d" % i, t)) for i, (r, t) in enumerate(outs)]
    proj = gandiva.make_projector(cases.Q1_SCHEMA, exprs, None, mode)
    st = devmem.stream()
    # the batch, resident on the device
    keep, cols = [], []
    for i in range(batch.num_columns):
        bufs = batch.column(i).buffers()
        w = cases.Q1_SCHEMA.field(i).type.bit_width // 8
        vals = devmem.DevBuf(n * w + 16, np.uint8)
        vld = devmem.DevBuf((n + 31) // 32 * 4 + 8, np.uint8, fill=0)
        gandiva.memcpy_htod(0, vals.ptr, np.frombuffer(bufs[1], dtype=np.uint8)[: n * w])
        gandiva.memcpy_htod(0, vld.ptr, np.frombuffer(bufs[0], dtype=np.uint8)[: (n + 7) // 8])
        keep += [vals, vld]
        cols.append((vld.ptr, vals.ptr, 0, 0))
    sel = devmem.DevBuf(n + 8, npt, fill=0)
    cnt = devmem.DevBuf(1, np.int64, fill=0)
    out_bufs = []
    for _, t in outs:
        v = devmem.DevBuf(n * (t.bit_width // 8) + 16, np.uint8, fill=0xEE)
        vl = devmem.DevBuf((n + 31) // 32 + 2, np.int32, fill=0)
        out_bufs.append((vl, v))
    launches = gandiva.launch_count()
    # ---- the chain: two enqueues, no synchronisation, no count on the host in between
    filt.evaluate_device(n, cols, sel.ptr, n, mode, st, cnt.ptr, sync=False)
    proj.evaluate_device(n, cols, [(vl.ptr, v.ptr) for vl, v in out_bufs], st, selection=(sel.ptr, n, cnt.ptr))
    assert gandiva.launch_count() - launches == 2
    proj.sync(st)
    count = filt.sync(st)
    # ---- against the oracle
    want_idx = oracle.filter_indices(cond, batch, threads=4)
    assert count == len(want_idx) == int(cnt.numpy()[0])
    assert np.array_equal(sel.numpy()[:count].astype(np.uint64), want_idx)
    want = oracle.project([r for r, _ in outs], [t for _, t in outs], batch, selection=want_idx.astype(np.int64), threads=4)
    for k, (_, t) in enumerate(outs):
        w = t.bit_width // 8
        vals = out_bufs[k][1].numpy()
        got = pa.Array.from_buffers(t, count, [pa.py_buffer(out_bufs[k][0].numpy().view(np.uint8)),
                                               pa.py_buffer(vals[: max(count * w, 1)].copy())])
        assert_arrays_match(got, want[k], "chained output %d (n=%d, %s)" % (k, n, mode))
        # rows past the device-side count were not written
        assert (vals[count * w: n * w] == 0xEE).all(), "output %d written past the selection count" % k


def test_raising_arguments_on_the_gpu(gandiva, oracle):
    """castVARCHAR(x, n < 0) and locate(.., start < 1) raise ExecutionError from the kernel, with the
    reference's messages; rows with a NULL argument, and rows in an if/else branch that is not taken, do not."""
    b = gandiva.TreeExprBuilder()
    S, L, I, B = pa.string(), pa.int64(), pa.int32(), pa.bool_()
    schema = pa.schema([("s", S), ("n", L), ("p", I)])
    s, n, p = cases.F(b, "s", S), cases.F(b, "n", L), cases.F(b, "p", I)
    fn = b.make_function
    cv = fn("char_length", [fn("castVARCHAR", [s, n], S)], I)
    loc = fn("locate", [b.make_literal("a", S), s, p], I)
    guarded = b.make_if(fn("greater_than_or_equal_to", [n, b.make_literal(0, L)], B), cv, b.make_literal(-1, I), I)
    proj = gandiva.make_projector(schema, [b.make_expression(cv, pa.field("a", I)), b.make_expression(loc, pa.field("b", I))], None)
    ok = pa.RecordBatch.from_arrays([pa.array(["banana", "x", None, "abc"] * 500), pa.array([3, 0, -5, None] * 500, L),
                                     pa.array([1, 2, 0, None] * 500, I)], schema=schema)
    got = proj.evaluate(ok)
    want = oracle.project([cv, loc], [I, I], ok)
    assert_arrays_match(got[0], want[0], "castVARCHAR with NULL rows")
    assert_arrays_match(got[1], want[1], "locate with NULL rows")
    bad_len = pa.RecordBatch.from_arrays([pa.array(["banana", "x"] * 70), pa.array([3, 1] * 69 + [3, -1], L),
                                          pa.array([1, 1] * 70, I)], schema=schema)
    with pytest.raises(gandiva.GandivaError, match="ExecutionError: Output buffer length can't be negative"):
        proj.evaluate(bad_len)
    bad_start = pa.RecordBatch.from_arrays([pa.array(["banana", "x"] * 70), pa.array([3, 1] * 70, L),
                                            pa.array([1, 1] * 69 + [0, 1], I)], schema=schema)
    with pytest.raises(gandiva.GandivaError, match="ExecutionError: Start position must be greater than 0"):
        proj.evaluate(bad_start)
    pg = gandiva.make_projector(schema, [b.make_expression(guarded, pa.field("g", I))], None)
    g, = pg.evaluate(bad_len)
    w, = oracle.project([guarded], [I], bad_len)
    assert_arrays_match(g, w, "guarded castVARCHAR")
    # as a Filter: the key-driven / hoisted shortcuts are off for conditions that can raise
    cond = fn("greater_than", [cv, b.make_literal(2, I)], B)
    f = gandiva.make_filter(schema, b.make_condition(cond))
    assert np.array_equal(f.evaluate(ok).to_array().to_numpy().astype(np.uint64), oracle.filter_indices(cond, ok))
    with pytest.raises(gandiva.GandivaError, match="can't be negative"):
        f.evaluate(bad_len)


def test_factorial_and_bround_on_the_gpu(gandiva, oracle):
    """factorial raises outside 0..20 with the reference's messages; bround rounds half to even."""
    b = gandiva.TreeExprBuilder()
    L, D = pa.int64(), pa.float64()
    schema = pa.schema([("l", L), ("d", D)])
    fact = b.make_function("factorial", [cases.F(b, "l", L)], L)
    br = b.make_function("bround", [cases.F(b, "d", D)], D)
    p = gandiva.make_projector(schema, [b.make_expression(fact, pa.field("f", L)), b.make_expression(br, pa.field("r", D))], None)
    ok = pa.RecordBatch.from_arrays([pa.array([0, 1, 5, 20, None] * 40, L), pa.array([0.5, 1.5, 2.5, -0.5, None] * 40, D)], schema=schema)
    got = p.evaluate(ok)
    want = oracle.project([fact, br], [L, D], ok)
    assert_arrays_match(got[0], want[0], "factorial")
    assert_arrays_match(got[1], want[1], "bround")
    assert got[1].to_pylist()[:4] == [0.0, 2.0, 2.0, -0.0]
    for bad, msg in (([3, -1], "Factorial of negative number not exist!"), ([21, 2], "Factorial of number greater than 20 not supported!")):
        batch = pa.RecordBatch.from_arrays([pa.array(bad * 50, L), pa.array([1.0, 2.0] * 50, D)], schema=schema)
        with pytest.raises(gandiva.GandivaError, match="ExecutionError: " + msg):
            p.evaluate(batch)
