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


def test_initcap_views_and_consumers(gandiva, oracle):
    """initcap is a lazy case map (it looks one byte back): projected, under concat / if / rtrim / upper / lower,
    over substr / trim, and read by like / equal / starts_with / length in a Filter — bit-exact against the oracle;
    a function that would move the start of the view AFTER initcap is refused at Make()."""
    S, B, I32, I64 = pa.string(), pa.bool_(), pa.int32(), pa.int64()
    schema = pa.schema([("s", S), ("u", S), ("p", B)])
    rng = np.random.default_rng(21)
    alphabet = list("abcdXYZ019 _-.,'") + ["é", "日"]
    n = 3000
    mk = lambda: pa.array([None if rng.random() < 0.1 else "".join(rng.choice(alphabet, size=int(rng.integers(0, 30))))
                           for _ in range(n)], S)
    batch = pa.RecordBatch.from_arrays([mk(), mk(), pa.array(rng.random(n) < 0.5, mask=rng.random(n) < 0.1)], schema=schema)
    b = gandiva.TreeExprBuilder()
    f = {x.name: b.make_field(x) for x in schema}
    fn = b.make_function
    ic = lambda x: fn("initcap", [x], S)
    roots = [(ic(f["s"]), S),
             (ic(fn("substr", [f["s"], b.make_literal(3, I64), b.make_literal(12, I64)], S)), S),
             (ic(fn("upper", [f["s"]], S)), S),
             (fn("upper", [ic(f["s"])], S), S),
             (fn("lower", [ic(f["s"])], S), S),
             (fn("rtrim", [ic(f["s"])], S), S),
             (ic(fn("btrim", [f["s"]], S)), S),
             (fn("concat", [ic(f["s"]), b.make_literal("|", S), ic(f["u"])], S), S),
             (b.make_if(f["p"], ic(f["s"]), f["u"], S), S),
             (fn("char_length", [ic(f["s"])], I32), I32)]
    for k, (root, t) in enumerate(roots):
        p = gandiva.make_projector(schema, [b.make_expression(root, pa.field("o", t))], None)
        got = p.evaluate(batch)[0]
        want = oracle.project([root], [t], batch)[0]
        assert_arrays_match(got, want, "initcap projection %d" % k)
    conds = [fn("like", [ic(f["s"]), b.make_literal("%Ab%", S)], B),
             fn("equal", [ic(f["s"]), ic(f["u"])], B),
             fn("starts_with", [ic(f["s"]), b.make_literal("A", S)], B),
             fn("ends_with", [ic(f["s"]), b.make_literal("b", S)], B),
             fn("less_than", [ic(f["s"]), f["u"]], B)]
    for k, cond in enumerate(conds):
        flt = gandiva.make_filter(schema, b.make_condition(cond))
        got = flt.evaluate(batch).to_array().to_numpy().astype(np.uint64)
        want = oracle.filter_indices(cond, batch)
        assert np.array_equal(got, want), "initcap filter %d" % k
    for bad in (fn("substr", [ic(f["s"]), b.make_literal(2, I64)], S), fn("ltrim", [ic(f["s"])], S),
                fn("reverse", [ic(f["s"])], S), fn("right", [ic(f["s"]), b.make_literal(3, I32)], S)):
        with pytest.raises(pa.ArrowNotImplementedError, match="initcap"):
            gandiva.make_projector(schema, [b.make_expression(bad, pa.field("o", S))], None)


def test_date_part_aliases_on_the_device(gandiva, oracle):
    ts, d64, I64 = pa.timestamp("ms"), pa.date64(), pa.int64()
    schema = pa.schema([("t", ts), ("d", d64)])
    rng = np.random.default_rng(4)
    n = 5000
    batch = pa.RecordBatch.from_arrays(
        [pa.array(rng.integers(-10**13, 10**13, n).astype(np.int64), ts, mask=rng.random(n) < 0.1),
         pa.array(rng.integers(-10**5, 10**5, n).astype(np.int64) * 86400000, d64, mask=rng.random(n) < 0.1)], schema=schema)
    b = gandiva.TreeExprBuilder()
    roots = []
    for name in ("year", "month", "day", "dayofmonth", "hour", "minute", "second", "dayofyear", "dayofweek", "quarter",
                 "weekofyear", "yearweek"):
        roots.append(b.make_function(name, [b.make_field(schema.field(0))], I64))
        roots.append(b.make_function(name, [b.make_field(schema.field(1))], I64))
    p = gandiva.make_projector(schema, [b.make_expression(r, pa.field("o%d" % k, I64)) for k, r in enumerate(roots)], None)
    got = p.evaluate(batch)
    want = oracle.project(roots, [I64] * len(roots), batch)
    for k, (gv, wv) in enumerate(zip(got, want)):
        assert_arrays_match(gv, wv, "alias output %d" % k)


def test_to_date_with_format(gandiva, oracle):
    """to_date(text, format literal [, suppress]): the format is compiled at Make() into a program the device
    function interprets; bit-exact against the oracle (itself refereed against libc's strptime), raising on the first
    text that does not parse unless errors are suppressed (NULL then)."""
    S, D64, I32 = pa.string(), pa.date64(), pa.int32()
    schema = pa.schema([("s", S)])
    rng = np.random.default_rng(23)
    b = gandiva.TreeExprBuilder()
    fs = b.make_field(schema.field(0))
    mons = ["Jan", "feb", "MARCH", "April", "may", "June", "jul", "AUGUST", "Sep", "october", "Nov", "DEC"]
    n = 4000

    def texts(kind):
        out = []
        for _ in range(n):
            y, m, d = int(rng.integers(1, 9999)), int(rng.integers(1, 13)), int(rng.integers(1, 29))
            if kind == 0:
                t = "%04d-%02d-%02d" % (y, m, d)
            elif kind == 1:
                t = "%d %s %d %02d:%02d" % (d, mons[m - 1], y, int(rng.integers(0, 24)), int(rng.integers(0, 60)))
            else:
                t = "%02d/%02d/%02d" % (m, d, y % 100)
            r = rng.random()
            if r < 0.05:
                t = t[: int(rng.integers(0, len(t)))]
            elif r < 0.10:
                t += "Z"
            elif r < 0.13:
                t = None
            out.append(t)
        return pa.array(out, S)
    for kind, fmt in enumerate(["YYYY-MM-DD", "DD MON YYYY HH24:MI", "MM/DD/YY"]):
        batch = pa.RecordBatch.from_arrays([texts(kind)], schema=schema)
        lenient = b.make_function("to_date", [fs, b.make_literal(fmt, S), b.make_literal(1, I32)], D64)
        p = gandiva.make_projector(schema, [b.make_expression(lenient, pa.field("d", D64))], None)
        got = p.evaluate(batch)[0]
        want = oracle.project([lenient], [D64], batch)[0]
        assert_arrays_match(got, want, "to_date %s" % fmt)
        assert 0 < want.null_count < n // 3
        strict = b.make_function("to_date", [fs, b.make_literal(fmt, S)], D64)
        ps = gandiva.make_projector(schema, [b.make_expression(strict, pa.field("d", D64))], None)
        with pytest.raises(gandiva.GandivaError, match="Error parsing value"):
            ps.evaluate(batch)
        clean = batch.filter(pa.compute.is_valid(want))
        assert_arrays_match(ps.evaluate(clean)[0], oracle.project([strict], [D64], clean)[0], "strict to_date %s" % fmt)
    with pytest.raises(pa.ArrowNotImplementedError, match="DDD"):
        gandiva.make_projector(schema, [b.make_expression(b.make_function("to_date", [fs, b.make_literal("YYYY-DDD", S)], D64),
                                                          pa.field("d", D64))], None)
    with pytest.raises(gandiva.GandivaError, match="requires a literal"):
        gandiva.make_projector(schema, [b.make_expression(b.make_function("to_date", [fs, fs], D64), pa.field("d", D64))], None)
