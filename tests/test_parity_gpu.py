"""Parity tests proper: the CUDA path, called through the C-ABI (ctypes mirror), against the
CPU oracle on the same seeded inputs.  Bit-exact for integer / decimal / string / date /
bool outputs and validity; float outputs within 0 ULP (BASELINE.json allows 1; kernels are
compiled with --fmad=false so the IEEE sequence is identical)."""
import numpy as np
import pyarrow as pa
import pytest

import cases
import devmem
from helpers import assert_arrays_match

pytestmark = pytest.mark.gpu

PROJECT_CASES = cases.all_project_cases()
FILTER_CASES = cases.all_filter_cases()


def _needs_small(name):
    return name.startswith(("divide", "mod", "in_", "and_short"))


def run_both(gandiva, oracle, build, n, seed, null_prob=0.15, offset=0, cfg=None):
    b = gandiva.TreeExprBuilder()
    schema, outs, kind = build(b)
    batch = cases.random_batch(schema, n, seed, null_prob, offset, small=_needs_small(build.__name__))
    exprs = [b.make_expression(r, pa.field("o%d" % i, t)) for i, (r, t) in enumerate(outs)]
    p = gandiva.make_projector(schema, exprs, None, "NONE", cfg)
    got = p.evaluate(batch)
    want = oracle.project([r for r, _ in outs], [t for _, t in outs], batch, threads=4)
    return got, want


@pytest.mark.parametrize("case", PROJECT_CASES, ids=[c.__name__ for c in PROJECT_CASES])
def test_project_parity(case, gandiva, oracle):
    for n, offset, seed in [(1, 0, 1), (33, 0, 2), (1000, 3, 3), (20011, 13, 4)]:
        got, want = run_both(gandiva, oracle, case, n, seed, offset=offset)
        for i, (g, w) in enumerate(zip(got, want)):
            assert_arrays_match(g, w, "%s n=%d out=%d" % (case.__name__, n, i))


@pytest.mark.parametrize("n", [1, 31, 32, 33, 255, 256, 257, 4095, 4096, 4097, 100003])
def test_project_sizes_no_nulls(n, gandiva, oracle):
    """Tail handling at warp / tile boundaries, absent validity buffers (buffers[0] == NULL)."""
    got, want = run_both(gandiva, oracle, cases.case_arith("add", pa.int32()), n, seed=n, null_prob=0.0)
    assert_arrays_match(got[0], want[0], "add int32 n=%d" % n)
    assert got[0].null_count == 0


@pytest.mark.parametrize("rpt", [1, 2, 4, 8, 16])
@pytest.mark.parametrize("bt", [128, 256, 512])
def test_project_tuning_knobs(rpt, bt, gandiva, oracle):
    cfg = gandiva.Configuration(rows_per_thread=rpt, block_threads=bt)
    got, want = run_both(gandiva, oracle, cases.case_if_else, 50021, seed=7, cfg=cfg)
    assert_arrays_match(got[0], want[0], "if_else rpt=%d bt=%d" % (rpt, bt))


@pytest.mark.parametrize("n", [255, 256, 257, 511, 512, 513, 767, 768, 769, 1024, 5000, 100003, 1_000_003])
@pytest.mark.parametrize("null_prob", [0.0, 0.2])
def test_project_tma_loader_sizes(n, null_prob, gandiva, oracle):
    """TMA bulk loader (loader=2): staged full tiles + direct tail, at tile boundaries; array
    offsets make every column pointer misaligned by a different amount."""
    cfg = gandiva.Configuration(rows_per_thread=2, block_threads=256, loader=2)
    for offset in (0, 3):
        got, want = run_both(gandiva, oracle, cases.case_if_else, n, seed=n + offset, null_prob=null_prob,
                             offset=offset, cfg=cfg)
        assert_arrays_match(got[0], want[0], "tma if_else n=%d off=%d" % (n, offset))


@pytest.mark.parametrize("rpt,bt,stages", [(1, 128, 2), (1, 256, 3), (2, 256, 2), (2, 256, 4), (4, 128, 3),
                                           (4, 256, 2), (2, 512, 2), (8, 128, 2)])
def test_project_tma_loader_knobs(rpt, bt, stages, gandiva, oracle):
    cfg = gandiva.Configuration(rows_per_thread=rpt, block_threads=bt, loader=2, stages=stages)
    b = gandiva.TreeExprBuilder()
    outs = cases.q1_outputs(b)
    batch = cases.q1_batch(150_001, seed=rpt * 1000 + bt + stages, null_permille=20)
    exprs = [b.make_expression(r, pa.field("o%d" % i, t)) for i, (r, t) in enumerate(outs)]
    p = gandiva.make_projector(cases.Q1_SCHEMA, exprs, None, "NONE", cfg)
    got = p.evaluate(batch)
    assert p.kernel_info["staged"] == 1
    want = oracle.project([r for r, _ in outs], [t for _, t in outs], batch, threads=4)
    for i, (g, w) in enumerate(zip(got, want)):
        assert_arrays_match(g, w, "q1 tma out %d" % i)


@pytest.mark.parametrize("case", PROJECT_CASES[::3], ids=[c.__name__ for c in PROJECT_CASES[::3]])
def test_project_tma_loader_cases(case, gandiva, oracle):
    """Same cases as test_project_parity through the staged loader (falls back to the direct
    path by itself when a column is bool / varlen)."""
    cfg = gandiva.Configuration(loader=2)
    got, want = run_both(gandiva, oracle, case, 20011, 4, offset=13, cfg=cfg)
    for i, (g, w) in enumerate(zip(got, want)):
        assert_arrays_match(g, w, "%s tma out=%d" % (case.__name__, i))


LIKE_SCAN_CASES = cases.all_like_scan_cases()


@pytest.mark.parametrize("case", LIKE_SCAN_CASES, ids=[c.__name__ for c in LIKE_SCAN_CASES])
def test_like_cooperative_scan(case, gandiva, oracle):
    """LIKE through the warp-cooperative scan of the staged bytes (hit list + per-row chain),
    against the oracle: plain rows, rows that overflow the hit list (per-lane fallback), rows
    longer than the shared-memory stage, non-ASCII rows, sliced arrays."""
    b = gandiva.TreeExprBuilder()
    schema, outs, _ = case(b)
    expr = b.make_expression(outs[0][0], pa.field("o", pa.bool_()))
    p = gandiva.make_projector(schema, [expr], None)
    assert "gdv_likeh_" in p.llvm_ir
    p_lane = gandiva.make_projector(schema, [expr], None, "NONE", gandiva.Configuration(string_scan=1))
    for n, seed, offset, dense, long_rows in [(64, 1, 0, False, False), (5000, 2, 0, False, False),
                                              (20011, 3, 7, True, False), (9000, 4, 1, False, True),
                                              (4097, 5, 3, True, True)]:
        batch = cases.like_scan_batch(n, seed, offset=offset, dense=dense, long_rows=long_rows)
        want, = oracle.project([outs[0][0]], [pa.bool_()], batch, threads=4)
        got, = p.evaluate(batch)
        assert_arrays_match(got, want, "%s n=%d" % (case.__name__, n))
        got_lane, = p_lane.evaluate(batch)
        assert_arrays_match(got_lane, want, "%s per-lane n=%d" % (case.__name__, n))


@pytest.mark.parametrize("bt,rpt", [(256, 2), (512, 2), (1024, 2), (512, 4), (256, 1)])
def test_like_scan_filter(bt, rpt, gandiva, oracle):
    """The config-4 condition as a Filter on l_comment-like rows through the ROW-DRIVEN string kernel
    (string_scan bit 2; the key-driven default is covered by test_filter_kernels_gpu.py): cooperative
    scan + cp.async stages at several block sizes and rows/thread."""
    b = gandiva.TreeExprBuilder()
    cond = cases.comment_condition(b)
    f = gandiva.make_filter(cases.COMMENT_SCHEMA, b.make_condition(cond),
                            gandiva.Configuration(rows_per_thread=rpt, block_threads=bt, string_scan=4))
    assert "gdv_likeh_" in f.llvm_ir
    for n in (70_001, 300_000):
        batch = cases.comment_batch(n, seed=bt + rpt)
        sel = f.evaluate(batch)
        want = oracle.filter_indices(cond, batch, threads=4)
        assert sel.num_slots == len(want) and len(want) > 0
        assert np.array_equal(sel.to_array().to_numpy().astype(np.uint64), want)


def _filter_both(gandiva, oracle, build, batch, dtype="int32", cfg=None):
    b = gandiva.TreeExprBuilder()
    schema, outs, kind = build(b)
    f = gandiva.make_filter(schema, b.make_condition(outs[0][0]), cfg)
    sel = f.evaluate(batch, None, dtype)
    want = oracle.filter_indices(outs[0][0], batch, threads=4)
    return sel, want


@pytest.mark.parametrize("case", FILTER_CASES, ids=[c.__name__ for c in FILTER_CASES])
@pytest.mark.parametrize("n", [1, 32, 33, 2047, 2048, 2049, 65536, 300007])
def test_filter_parity(case, n, gandiva, oracle):
    b = gandiva.TreeExprBuilder()
    schema, _, _ = case(b)
    if case is cases.case_q6_filter:
        batch = cases.q6_batch(n, seed=42, null_permille=10 if n % 2 else 0)
    else:
        batch = cases.random_batch(schema, n, seed=n, null_prob=0.1, offset=5 if n > 100 else 0, small=True)
    sel, want = _filter_both(gandiva, oracle, case, batch)
    got = sel.to_array().to_numpy()
    assert got.dtype == np.uint32
    assert sel.num_slots == len(want), "%s n=%d count %d != %d" % (case.__name__, n, sel.num_slots, len(want))
    assert np.array_equal(got.astype(np.uint64), want)


@pytest.mark.parametrize("dtype,npt", [("int16", np.uint16), ("int32", np.uint32), ("int64", np.uint64)])
def test_filter_index_widths(dtype, npt, gandiva, oracle):
    batch = cases.q6_batch(60000, seed=7)
    sel, want = _filter_both(gandiva, oracle, cases.case_q6_filter, batch, dtype)
    got = sel.to_array().to_numpy()
    assert got.dtype == npt
    assert np.array_equal(got.astype(np.uint64), want)


@pytest.mark.parametrize("rpt,bt", [(1, 128), (2, 256), (4, 256), (8, 256), (8, 512), (16, 1024)])
def test_filter_tuning_knobs(rpt, bt, gandiva, oracle):
    cfg = gandiva.Configuration(rows_per_thread=rpt, block_threads=bt)
    batch = cases.q6_batch(777777, seed=9, null_permille=20)
    sel, want = _filter_both(gandiva, oracle, cases.case_q6_filter, batch, cfg=cfg)
    assert np.array_equal(sel.to_array().to_numpy().astype(np.uint64), want)


def test_filter_selection_too_small(gandiva):
    b = gandiva.TreeExprBuilder()
    schema, outs, _ = cases.case_filter_all(b)
    f = gandiva.make_filter(schema, b.make_condition(outs[0][0]))
    batch = cases.random_batch(schema, 70000, seed=1)
    with pytest.raises(pa.ArrowInvalid):
        f.evaluate(batch, None, "int16")  # > 65536 rows cannot be indexed by uint16


@pytest.mark.parametrize("mode,npt", [("UINT16", np.uint16), ("UINT32", np.uint32), ("UINT64", np.uint64)])
def test_project_with_selection_vector(mode, npt, gandiva, oracle):
    b = gandiva.TreeExprBuilder()
    schema, outs, _ = cases.case_if_else(b)
    n = 50000
    batch = cases.random_batch(schema, n, seed=21, null_prob=0.2, offset=7)
    rng = np.random.default_rng(5)
    idx = np.sort(rng.choice(n, 12345, replace=False)).astype(npt)
    exprs = [b.make_expression(r, pa.field("o%d" % i, t)) for i, (r, t) in enumerate(outs)]
    p = gandiva.make_projector(schema, exprs, None, mode)
    sel = gandiva.SelectionVector(idx, len(idx), gandiva._SEL_MODE[mode])
    got = p.evaluate(batch, sel)
    want = oracle.project([r for r, _ in outs], [t for _, t in outs], batch, selection=idx.astype(np.int64))
    assert len(got[0]) == len(idx)
    assert_arrays_match(got[0], want[0], "selection %s" % mode)


def test_q1_projector_on_lineitem(gandiva, oracle):
    """Eight outputs from one fused kernel (config 3 shape) on the synthetic lineitem columns."""
    b = gandiva.TreeExprBuilder()
    outs = cases.q1_outputs(b)
    batch = cases.q1_batch(200_003, seed=42, null_permille=20)
    exprs = [b.make_expression(r, pa.field("o%d" % i, t)) for i, (r, t) in enumerate(outs)]
    p = gandiva.make_projector(cases.Q1_SCHEMA, exprs, None)
    got = p.evaluate(batch)
    want = oracle.project([r for r, _ in outs], [t for _, t in outs], batch, threads=4)
    for i, (g, w) in enumerate(zip(got, want)):
        assert_arrays_match(g, w, "q1 out %d" % i)
    assert got[0].null_count > 0 and got[7].null_count > got[4].null_count


@pytest.mark.parametrize("t", [pa.int32(), pa.int64(), pa.float64()], ids=str)
def test_divide_by_zero_is_execution_error(t, gandiva, oracle):
    b = gandiva.TreeExprBuilder()
    schema = pa.schema([("a", t), ("b", t)])
    root = b.make_function("divide", [cases.F(b, "a", t), cases.F(b, "b", t)], t)
    p = gandiva.make_projector(schema, [b.make_expression(root, pa.field("r", t))], None)
    av = pa.array([1, 2, 3, 4], t)
    batch = pa.RecordBatch.from_arrays([av, pa.array([1, 0, 2, 1], t)], schema=schema)
    with pytest.raises(gandiva.GandivaError, match="ExecutionError: divide by zero error"):
        p.evaluate(batch)
    with pytest.raises(Exception, match="divide by zero"):
        oracle.project([root], [t], batch)
    # the same projector keeps working afterwards, and a NULL divisor does not raise
    batch = pa.RecordBatch.from_arrays([av, pa.array([1, None, 2, 1], t)], schema=schema)
    got, = p.evaluate(batch)
    want, = oracle.project([root], [t], batch)
    assert_arrays_match(got, want, "divide with null divisor")


def test_decimal_divide_by_zero_is_execution_error(gandiva, oracle):
    import decimal
    D = decimal.Decimal
    for build in (cases.case_decimal_divide(15, 2, 15, 2, guarded=False),):
        b = gandiva.TreeExprBuilder()
        schema, outs, _ = build(b)
        p = gandiva.make_projector(schema, [b.make_expression(outs[0][0], pa.field("r", outs[0][1]))], None)
        batch = pa.RecordBatch.from_arrays([pa.array([D("1.00"), D("2.00"), D("7.25")], schema.field(0).type),
                                            pa.array([D("3.00"), D("0.00"), D("-0.50")], schema.field(1).type)],
                                           schema=schema)
        with pytest.raises(gandiva.GandivaError, match="ExecutionError: divide by zero error"):
            p.evaluate(batch)
        batch = pa.RecordBatch.from_arrays([batch.column(0), pa.array([D("3.00"), None, D("-0.50")], schema.field(1).type)],
                                           schema=schema)
        got, = p.evaluate(batch)
        want, = oracle.project([outs[0][0]], [outs[0][1]], batch)
        assert_arrays_match(got, want, "decimal divide with a null divisor")
        assert got.to_pylist()[2] == D("-14.5")


@pytest.mark.parametrize("n,offset", [(1, 0), (40, 0), (513, 2), (30011, 5)])
def test_string_outputs(n, offset, gandiva, oracle):
    """utf8 outputs (sizing pass, tile scan, write pass) next to a fixed-width output, small
    substr / castVARCHAR arguments, sliced inputs; plus the sizing entry point on its own."""
    b = gandiva.TreeExprBuilder()
    schema, outs, _ = cases.case_string_outputs(b)
    batch = cases.random_batch(schema, n, seed=n, null_prob=0.15, offset=offset, small=True)
    exprs = [b.make_expression(r, pa.field("o%d" % i, t)) for i, (r, t) in enumerate(outs)]
    p = gandiva.make_projector(schema, exprs, None)
    got = p.evaluate(batch)
    want = oracle.project([r for r, _ in outs], [t for _, t in outs], batch)
    for i, (g, w) in enumerate(zip(got, want)):
        assert_arrays_match(g, w, "string outputs n=%d out=%d" % (n, i))
        if pa.types.is_string(g.type):
            g.validate(full=True)
    # a projector with only string outputs, and with a selection vector
    p2 = gandiva.make_projector(schema, exprs[1:3], None, "UINT32")
    idx = np.sort(np.random.default_rng(n).choice(n, max(1, n // 3), replace=False)).astype(np.uint32)
    sel = gandiva.SelectionVector(idx, len(idx), gandiva._SEL_MODE["UINT32"])
    got2 = p2.evaluate(batch, sel)
    want2 = oracle.project([r for r, _ in outs[1:3]], [t for _, t in outs[1:3]], batch, selection=idx.astype(np.int64))
    for g, w in zip(got2, want2):
        assert_arrays_match(g, w, "string outputs with selection n=%d" % n)


def test_empty_batch_rejected(gandiva):
    b = gandiva.TreeExprBuilder()
    schema, outs, _ = cases.case_arith("add", pa.int32())(b)
    p = gandiva.make_projector(schema, [b.make_expression(outs[0][0], pa.field("r", pa.int32()))], None)
    batch = cases.random_batch(schema, 0, seed=1)
    with pytest.raises(pa.ArrowInvalid, match="non-empty"):
        p.evaluate(batch)


def test_generator_matches_cpu_twin(gandiva, oracle):
    """Device lineitem generator == oracle/lineitem.h for every column kind."""
    n = 100003
    st = devmem.stream()
    for kind, npdt in [(0, np.int32), (1, np.float64), (2, np.float64), (3, np.int64), (7, np.float64),
                       (8, np.float64), (9, np.int32)]:
        vals = devmem.DevBuf(n, npdt, fill=0)
        vld = devmem.DevBuf((n + 31) // 32, np.int32, fill=0)
        gandiva.generate_lineitem(0, kind, 42, 1000, n, vals.ptr, vld.ptr, 15, st)
        devmem.synchronize()
        cv, cvld = oracle.generate_lineitem(kind, 42, 1000, n, 15)
        assert np.array_equal(vals.numpy(), cv), kind
        gbits = np.unpackbits(vld.numpy().view(np.uint8), bitorder="little")[:n]
        cbits = np.unpackbits(cvld, bitorder="little")[:n]
        assert np.array_equal(gbits, cbits), kind
    for kind in (4, 5, 6):
        vals = devmem.DevBuf((n, 2), np.int64, fill=0)
        gandiva.generate_lineitem(0, kind, 42, 0, n, vals.ptr, 0, 0, st)
        devmem.synchronize()
        cv, _ = oracle.generate_lineitem(kind, 42, 0, n)
        assert np.array_equal(vals.numpy().view(np.uint64), cv), kind


def test_device_resident_filter_and_properties(gandiva, oracle):
    """Device-resident async API (the bench path): inputs generated in HBM, indices stay in HBM.
    Checked exactly against the oracle on the first rows and by size-independent properties
    (ascending, count == independent mask count, indices == nonzero(mask)) on all rows."""
    # >= 32M rows: exercises the large-batch kernel variant (1024-thread tiles); the simulator
    # takes a batch it finishes in seconds (that variant is reached there through block_threads)
    n = 300_007 if devmem.EMU else (32 << 20) + 3
    ship = devmem.DevBuf(n, np.int32)
    disc = devmem.DevBuf(n, np.float64)
    qty = devmem.DevBuf(n, np.float64)
    st = devmem.stream()
    for kind, tns in ((0, ship), (1, disc), (2, qty)):
        gandiva.generate_lineitem(0, kind, 42, 0, n, tns.ptr, 0, 0, st)
    b = gandiva.TreeExprBuilder()
    cfg = gandiva.Configuration(block_threads=1024) if devmem.EMU else None
    f = gandiva.make_filter(cases.Q6_SCHEMA, b.make_condition(cases.q6_condition(b)), cfg)
    out = devmem.DevBuf(n, np.int32)
    cnt = devmem.DevBuf(1, np.int64, fill=0)
    cols = [(0, ship.ptr, 0, 0), (0, disc.ptr, 0, 0), (0, qty.ptr, 0, 0)]
    for _ in range(3):  # repeated async launches on one stream reuse the look-back scratch
        f.evaluate_device(n, cols, out.ptr, n, "UINT32", st, cnt.ptr)
    count = f.sync(st)
    assert count == int(cnt.numpy()[0])
    if devmem.EMU:
        shipn, discn, qtyn = ship.numpy(), disc.numpy(), qty.numpy()
        mask = (shipn >= 8766) & (shipn < 9131) & (discn >= 0.05) & (discn <= 0.07) & (qtyn < 24)
        assert count == int(mask.sum())
        idx = out.numpy()[:count].astype(np.int64)
        assert bool((idx[1:] > idx[:-1]).all())
        assert np.array_equal(idx, np.nonzero(mask)[0])
    else:
        import torch
        mask = (ship.a >= 8766) & (ship.a < 9131) & (disc.a >= 0.05) & (disc.a <= 0.07) & (qty.a < 24)
        assert count == int(mask.sum().item())
        tidx = out.a[:count].to(torch.int64)
        assert bool((tidx[1:] > tidx[:-1]).all())
        assert torch.equal(tidx, torch.nonzero(mask).flatten())
        idx = tidx.cpu().numpy()
    # exact oracle parity on a prefix
    m = 200_000
    batch = cases.q6_batch(m, seed=42)
    assert np.array_equal(ship.numpy()[:m], batch.column(0).cast(pa.int32()).to_numpy(zero_copy_only=False))
    want = oracle.filter_indices(cases.q6_condition(b), batch, threads=4)
    got = idx[idx < m].astype(np.uint64)
    assert np.array_equal(got, want)


def test_device_resident_projector(gandiva, oracle):
    n = 300_001 if devmem.EMU else 3_000_001
    a = devmem.DevBuf(n, np.int32)
    bb = devmem.DevBuf(n, np.int32)
    av = devmem.DevBuf((n + 31) // 32, np.int32)
    bv = devmem.DevBuf((n + 31) // 32, np.int32)
    st = devmem.stream()
    gandiva.generate_lineitem(0, 9, 42, 0, n, a.ptr, av.ptr, 100, st)
    gandiva.generate_lineitem(0, 10, 42, 0, n, bb.ptr, bv.ptr, 100, st)
    bld = gandiva.TreeExprBuilder()
    t = pa.int32()
    schema = pa.schema([("a", t), ("b", t)])
    root = bld.make_function("add", [cases.F(bld, "a", t), cases.F(bld, "b", t)], t)
    p = gandiva.make_projector(schema, [bld.make_expression(root, pa.field("c", t))], None)
    out = devmem.DevBuf(n, np.int32)
    ov = devmem.DevBuf((n + 31) // 32, np.int32)
    p.evaluate_device(n, [(av.ptr, a.ptr, 0, 0), (bv.ptr, bb.ptr, 0, 0)], [(ov.ptr, out.ptr)], st)
    p.sync(st)
    ca, cav = oracle.generate_lineitem(9, 42, 0, n, 100, threads=4)
    cb, cbv = oracle.generate_lineitem(10, 42, 0, n, 100, threads=4)
    batch = pa.RecordBatch.from_arrays(
        [pa.Array.from_buffers(t, n, [pa.py_buffer(cav), pa.py_buffer(ca)]),
         pa.Array.from_buffers(t, n, [pa.py_buffer(cbv), pa.py_buffer(cb)])], schema=schema)
    want, = oracle.project([root], [t], batch, threads=4)
    got = pa.Array.from_buffers(t, n, [pa.py_buffer(ov.numpy()), pa.py_buffer(out.numpy())])
    assert_arrays_match(got, want, "device-resident add")


def test_kernels_were_launched(gandiva):
    assert gandiva.launch_count() > 0


def test_filter_bounded_selection_vector(gandiva, oracle):
    """GDV_SEL_BOUNDED: capacity smaller than the number of selected rows -> the count is still
    exact, the first max_slots indices are stored, nothing is written past the capacity."""
    n = 200_003 if devmem.EMU else 1_000_003
    st = devmem.stream()
    ship = devmem.DevBuf(n, np.int32)
    disc = devmem.DevBuf(n, np.float64)
    qty = devmem.DevBuf(n, np.float64)
    for kind, t in ((0, ship), (1, disc), (2, qty)):
        gandiva.generate_lineitem(0, kind, 42, 0, n, t.ptr, 0, 0, st)
    b = gandiva.TreeExprBuilder()
    f = gandiva.make_filter(cases.Q6_SCHEMA, b.make_condition(cases.q6_condition(b)))
    want = oracle.filter_indices(cases.q6_condition(b), cases.q6_batch(n, seed=42), threads=4)
    cols = [(0, ship.ptr, 0, 0), (0, disc.ptr, 0, 0), (0, qty.ptr, 0, 0)]
    for cap in (len(want) + 100, len(want) // 2, 7):
        out = devmem.DevBuf(cap + 64, np.int64, fill=-1)
        cnt = devmem.DevBuf(1, np.int64, fill=0)
        f.evaluate_device(n, cols, out.ptr, cap, "UINT64|BOUNDED", st, cnt.ptr, index_base=5)
        count = f.sync(st)
        assert count == len(want)
        k = min(cap, count)
        got = out.numpy()
        assert np.array_equal(got[:k].astype(np.uint64), want[:k] + 5)
        assert (got[max(cap, k):] == -1).all()
    with pytest.raises(pa.ArrowInvalid):   # without the flag the reference rule holds
        f.evaluate_device(n, cols, out.ptr, 7, "UINT64", st, cnt.ptr)


@pytest.mark.parametrize("waves", [1, 3])
def test_peer_selection_push(waves, gandiva):
    """Multi-GPU reassembly of the SelectionVector (needs >= 2 GPUs on the box); one run per rank, and the
    batch filtered in waves whose pushes hide under the next wave's filter kernel."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(root, "tests", "peer_push_worker.py"), "5000003", "5", str(waves)],
                         capture_output=True, text=True, timeout=600)
    assert "PEER_PUSH_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


def _int_strings(n, seed, bits):
    rng = np.random.default_rng(seed)
    lo, hi = -(1 << (bits - 1)), (1 << (bits - 1)) - 1
    vals = [int(v) for v in rng.integers(lo, hi, n, dtype=np.int64, endpoint=True)]
    vals[:4] = [lo, hi, 0, -1][:n]
    out = []
    for k, v in enumerate(vals):
        t = str(v) if k % 5 else ("+" + str(v) if v >= 0 else str(v))
        t = (" " * (k % 3)) + t + (" " * (k % 2))
        out.append(None if k % 11 == 7 else t)
    return out, vals


@pytest.mark.parametrize("fn_name,t,bits", [("castINT", pa.int32(), 32), ("castBIGINT", pa.int64(), 64)])
def test_cast_string_to_integer(fn_name, t, bits, gandiva, oracle):
    """castINT / castBIGINT(utf8): spaces, signs, the extreme values; strings that are not integers
    of the type raise an ExecutionError (in the kernel and in the oracle), a NULL row does not."""
    b = gandiva.TreeExprBuilder()
    S = pa.string()
    schema = pa.schema([("s", S)])
    root = b.make_function(fn_name, [cases.F(b, "s", S)], t)
    p = gandiva.make_projector(schema, [b.make_expression(root, pa.field("v", t))], None)
    for n in (1, 40, 3001):
        strs, vals = _int_strings(n, n, bits)
        batch = pa.RecordBatch.from_arrays([pa.array(strs, S)], schema=schema)
        got, = p.evaluate(batch)
        want, = oracle.project([root], [t], batch)
        assert_arrays_match(got, want, "%s n=%d" % (fn_name, n))
        assert got.to_pylist() == [None if s is None else v for s, v in zip(strs, vals)]
    too_big = str(1 << (bits - 1))
    for bad in ("12a", "", "  ", "-", "1 2", too_big, "-" + str((1 << (bits - 1)) + 1), "9" * 30, "１２"):
        batch = pa.RecordBatch.from_arrays([pa.array(["7", None, bad, "8"], S)], schema=schema)
        with pytest.raises(gandiva.GandivaError, match="ExecutionError: Failed to cast"):
            p.evaluate(batch)
        with pytest.raises(Exception, match="Failed to cast"):
            oracle.project([root], [t], batch)
    batch = pa.RecordBatch.from_arrays([pa.array(["7", None, "-8"], S)], schema=schema)
    assert p.evaluate(batch)[0].to_pylist() == [7, None, -8]


def _float_strings(n, seed):
    rng = np.random.default_rng(seed)
    out = ["0", "-0", "+0.0", "1", "-1.5", " 2.5 ", "1e0", "1E+2", "1.e2", ".5", "5.", "-.5e-1", "1e22", "1e23",
           "9007199254740993", "9007199254740992.5", "0.1", "0.3", "1e-38", "1.7976931348623157e308", "1e309",
           "-1e400", "4.9e-324", "2.2250738585072014e-308", "1e-400", "123456789012345678901234567890",
           "0.000000000000000000000000000000000000001", "3.4028235e38", "1.17549435e-38", "1e-45",
           "0000000000000000000000000001", "0.00000", "100000000000000000000000000000.00000000001e-29",
           "2.2250738585072011e-308", "2.2250738585072012e-308", "4.9406564584124654e-324", "2.4703282292062327e-324",
           "2.4703282292062328e-324", "7.4109846876186982e-324", "1.7976931348623158e308", "1.7976931348623159e308",
           "17976931348623158079e289", "8.98846567431158e307", "9007199254740992", "9007199254740994", "9007199254740995",
           "0.500000000000000166533453693773481063544750213623046875", "6.2230152778611417e-1",
           "8.5e-324", "1e-323", "3e-324", "2e-324", "5e-324", "1.0000000000000002", "1.00000000000000011102230246251565",
           "1.00000000000000033306690738754696", "9.5e-1", "4.35e0", "1e+400", "1e-99999999999", "1e99999999999",
           "0e99999999999", "7450580596923828125e-27", "14901161193847656250e-28", "1e-5", "5e-1", "123456789e-9"]
    while len(out) < n:
        style = int(rng.integers(0, 5))
        digits = "".join(str(int(d)) for d in rng.integers(0, 10, size=int(rng.integers(1, 22))))
        if style == 0:
            t = digits
        elif style == 1:
            cut = int(rng.integers(0, len(digits) + 1))
            t = digits[:cut] + "." + digits[cut:]
        elif style == 2:
            t = digits[:1] + "." + digits[1:] + "e%d" % int(rng.integers(-45, 46))
        elif style == 3:
            t = digits + "E%+d" % int(rng.integers(-330, 310))
        else:
            t = repr(abs(float(rng.standard_normal() * 10.0 ** int(rng.integers(-30, 30)))))
        if rng.integers(0, 2):
            t = "-" + t
        if rng.integers(0, 8) == 0:
            t = " " + t + "  "
        out.append(t)
    out = out[:n]
    for k in range(3, n, 11):
        out[k] = None
    return out


def _sig_digits(text):
    t = text.strip().lstrip("+-").lower().partition("e")[0].replace(".", "")
    return len(t.lstrip("0").rstrip("0"))


def test_cast_string_to_float(gandiva, oracle):
    """castFLOAT8 / castFLOAT4(utf8): kernel == oracle bit for bit; both == Python's correctly
    rounded float() over the whole double range (subnormals, overflow to inf, underflow to 0) when
    the text has <= 19 significant digits, within 1 ULP with more; malformed text raises in both."""
    b = gandiva.TreeExprBuilder()
    S, F8, F4 = pa.string(), pa.float64(), pa.float32()
    schema = pa.schema([("s", S)])
    root8 = b.make_function("castFLOAT8", [cases.F(b, "s", S)], F8)
    root4 = b.make_function("castFLOAT4", [cases.F(b, "s", S)], F4)
    p = gandiva.make_projector(schema, [b.make_expression(root8, pa.field("d", F8)),
                                        b.make_expression(root4, pa.field("f", F4))], None)
    for n, seed in ((1, 1), (33, 2), (2500, 3)):
        strs = _float_strings(n, seed)
        batch = pa.RecordBatch.from_arrays([pa.array(strs, S)], schema=schema)
        got8, got4 = p.evaluate(batch)
        want8, want4 = oracle.project([root8, root4], [F8, F4], batch)
        assert_arrays_match(got8, want8, "castFLOAT8 n=%d" % n)
        assert_arrays_match(got4, want4, "castFLOAT4 n=%d" % n)
        for text, v in zip(strs, got8.to_pylist()):
            if text is None:
                assert v is None
                continue
            ref = float(text)
            if _sig_digits(text) <= 19:
                assert np.float64(v).tobytes() == np.float64(ref).tobytes(), (text, v, ref)
            else:
                assert v == ref or v in (np.nextafter(ref, np.inf), np.nextafter(ref, -np.inf)), (text, v, ref)
    for bad in ("", " ", "-", ".", "e5", "1e", "1e+", "1.2.3", "1 2", "0x10", "nan", "inf", "1f", "--1", "１"):
        batch = pa.RecordBatch.from_arrays([pa.array(["7", None, bad, "8"], S)], schema=schema)
        with pytest.raises(gandiva.GandivaError, match="ExecutionError: Failed to cast the string to a float"):
            p.evaluate(batch)
        with pytest.raises(Exception, match="Failed to cast the string to a float"):
            oracle.project([root8], [F8], batch)


@pytest.mark.parametrize("chunk", range(6))
def test_regexp_matches(chunk, gandiva, oracle):
    """regexp_matches / regexp_like: the byte automaton built at Make() against the oracle's
    backtracking matcher over code points -- literals, classes, repetition, alternation, anchors,
    multi-byte text, NULLs; plain column, lower() view and substr() view as the subject."""
    b = gandiva.TreeExprBuilder()
    S, B, L = pa.string(), pa.bool_(), pa.int64()
    schema = pa.schema([("s", S)])
    s = cases.F(b, "s", S)
    pats = list(cases.REGEX_PATTERNS[chunk::6])
    while True:   # the random patterns may need more than 128 automaton positions: Make() names the first that does
        roots = [b.make_function("regexp_matches", [s, b.make_literal(p, S)], B) for p in pats]
        roots.append(b.make_function("regexp_like", [b.make_function("lower", [s], S), b.make_literal(pats[0], S)], B))
        roots.append(b.make_function("regexp_matches", [b.make_function("substr", [s, b.make_literal(2, L), b.make_literal(4, L)], S),
                                                        b.make_literal(pats[-1], S)], B))
        try:
            p = gandiva.make_projector(schema, [b.make_expression(r, pa.field("m%d" % i, B)) for i, r in enumerate(roots)], None)
            break
        except pa.ArrowNotImplementedError as e:
            too_big = [q for q in pats if "'%s' needs more than 128 automaton positions" % q in str(e)]
            assert len(too_big) >= 1, str(e)
            pats = [q for q in pats if q not in too_big]
    assert len(pats) >= 18, pats
    for n, seed in ((1, 1), (70, 2), (2000, 3 + chunk)):
        batch = pa.RecordBatch.from_arrays([pa.array(cases.regex_texts(n, seed), S)], schema=schema)
        got = p.evaluate(batch)
        want = oracle.project(roots, [B] * len(roots), batch)
        for i, (g, w) in enumerate(zip(got, want)):
            assert_arrays_match(g, w, "regexp %r n=%d" % (pats[i] if i < len(pats) else "view", n))
    cond = b.make_condition(roots[0])
    f = gandiva.make_filter(schema, cond)
    batch = pa.RecordBatch.from_arrays([pa.array(cases.regex_texts(3000, 40 + chunk), S)], schema=schema)
    sel = f.evaluate(batch, None).to_array().to_numpy()
    assert np.array_equal(sel.astype(np.uint64), oracle.filter_indices(roots[0], batch))


def test_regexp_pattern_errors(gandiva):
    b = gandiva.TreeExprBuilder()
    S, B = pa.string(), pa.bool_()
    schema = pa.schema([("s", S), ("t", S)])
    s, t = cases.F(b, "s", S), cases.F(b, "t", S)

    def make(pat_node):
        root = b.make_function("regexp_matches", [s, pat_node], B)
        return gandiva.make_projector(schema, [b.make_expression(root, pa.field("m", B))], None)
    for bad in ("(ab", "ab)", "[abc", "a**b(", "*a", "a\\", "[[:nope:]]", "\\x4"):
        with pytest.raises(Exception, match="regular expression"):
            make(b.make_literal(bad, S))
    for unsupported in ("\\bword\\b", "a(?i)bc", "(?=a)b", "a^b", "(a$)|b", "[α-ω]", "(a)\\1", "a{100}b{100}", "(\\w\\d\\s){50}", "\\pL", "\\xe9"):
        with pytest.raises(pa.ArrowNotImplementedError):
            make(b.make_literal(unsupported, S))
    with pytest.raises(Exception, match="requires a literal"):
        make(t)


@pytest.mark.parametrize("n", [1, 100, 4000])
def test_cast_float_to_string(n, gandiva, oracle):
    """castVARCHAR(float64 / float32, len): the kernel's fixed-point interval search for the shortest
    digits against the oracle's C-library search (which tests/test_oracle_vs_arrow.py pins to Python's
    repr); binade boundaries, subnormals, the layout thresholds 10^-3 and 10^7, truncation to len."""
    from test_oracle_vs_arrow import float_text_values
    b = gandiva.TreeExprBuilder()
    D, F4, S, L = pa.float64(), pa.float32(), pa.string(), pa.int64()
    schema = pa.schema([("d", D), ("f", F4)])
    vals = float_text_values(max(n // 3, 1), 17 + n)[-n:] if n < 100 else float_text_values(n // 3, 17 + n)
    with np.errstate(over="ignore"):
        f32 = np.array(vals, dtype=np.float64).astype(np.float32)
    nulls = np.arange(len(vals)) % 9 == 4
    batch = pa.RecordBatch.from_arrays([pa.array(vals, D, mask=nulls), pa.array(f32, F4, mask=nulls)], schema=schema)
    d, f = cases.F(b, "d", D), cases.F(b, "f", F4)
    fn = b.make_function
    roots = [fn("castVARCHAR", [d, b.make_literal(40, L)], S), fn("castVARCHAR", [f, b.make_literal(40, L)], S),
             fn("castVARCHAR", [d, b.make_literal(5, L)], S),
             fn("concat", [b.make_literal("v=", S), fn("castVARCHAR", [d, b.make_literal(30, L)], S)], S),
             fn("castFLOAT8", [fn("castVARCHAR", [d, b.make_literal(40, L)], S)], D)]
    types = [S, S, S, S, D]
    # the round trip text -> double only parses plain / exponent spellings: NaN and the infinities are left out of it
    finite = pa.RecordBatch.from_arrays([pa.array([v if np.isfinite(v) else 1.0 for v in vals], D, mask=nulls), batch.column(1)], schema=schema)
    p = gandiva.make_projector(schema, [b.make_expression(r, pa.field("o%d" % i, t)) for i, (r, t) in enumerate(zip(roots, types))], None)
    got = p.evaluate(finite)
    want = oracle.project(roots, types, finite, threads=4)
    for i, (g, w) in enumerate(zip(got, want)):
        assert_arrays_match(g, w, "castVARCHAR(float) out=%d n=%d" % (i, n))
    back = got[4].to_numpy(zero_copy_only=False)
    src = finite.column(0).to_numpy(zero_copy_only=False)
    ok = ~nulls[:len(src)]
    assert np.array_equal(back[ok].view(np.uint64), src[ok].view(np.uint64))   # shortest digits read back exactly
    p2 = gandiva.make_projector(schema, [b.make_expression(roots[0], pa.field("o", S))], None)
    assert_arrays_match(p2.evaluate(batch)[0], oracle.project(roots[:1], [S], batch)[0], "NaN / Infinity spellings")


def test_log_with_base(gandiva, oracle):
    """log(base, value) = ln(value) / ln(base), bit-exact with the oracle; base 1 raises like a division by zero."""
    b = gandiva.TreeExprBuilder()
    D = pa.float64()
    schema = pa.schema([("b", D), ("v", D)])
    root = b.make_function("log", [cases.F(b, "b", D), cases.F(b, "v", D)], D)
    p = gandiva.make_projector(schema, [b.make_expression(root, pa.field("r", D))], None)
    rng = np.random.default_rng(4)
    bases = np.concatenate([rng.uniform(1.5, 100, 500), [2.0, 10.0, 0.5, 1e-300, 7.0]])
    vals = np.concatenate([np.exp(rng.uniform(-300, 300, 500)), [8.0, 1000.0, 0.25, 1e300, 0.0]])
    batch = pa.RecordBatch.from_arrays([pa.array(bases, D), pa.array(vals, D)], schema=schema)
    got, = p.evaluate(batch)
    want, = oracle.project([root], [D], batch)
    assert_arrays_match(got, want, "log(base, value)")
    g = got.to_pylist()
    assert abs(g[500] - 3.0) < 1e-15 and abs(g[501] - 3.0) < 1e-15 and abs(g[502] - 2.0) < 1e-15 and g[504] == -np.inf
    bad = pa.RecordBatch.from_arrays([pa.array([2.0, 1.0], D), pa.array([4.0, 4.0], D)], schema=schema)
    with pytest.raises(gandiva.GandivaError, match="ExecutionError: divide by zero error"):
        p.evaluate(bad)
    with pytest.raises(Exception, match="divide by zero"):
        oracle.project([root], [D], bad)


def test_cast_string_to_boolean(gandiva, oracle):
    """castBIT / castBOOLEAN(utf8): kernel == oracle; anything but true / false / 1 / 0 raises in both."""
    b = gandiva.TreeExprBuilder()
    S, B = pa.string(), pa.bool_()
    schema = pa.schema([("s", S)])
    root = b.make_function("castBIT", [cases.F(b, "s", S)], B)
    root2 = b.make_function("castBOOLEAN", [b.make_function("btrim", [cases.F(b, "s", S)], S)], B)
    p = gandiva.make_projector(schema, [b.make_expression(root, pa.field("v", B)), b.make_expression(root2, pa.field("w", B))], None)
    words = ["true", " FALSE ", "1", "0", "True", None, "tRuE  ", "false", " 0"]
    for n in (1, 9, 1000):
        batch = pa.RecordBatch.from_arrays([pa.array([words[k % len(words)] for k in range(n)], S)], schema=schema)
        got = p.evaluate(batch)
        want = oracle.project([root, root2], [B, B], batch)
        assert_arrays_match(got[0], want[0], "castBIT")
        assert_arrays_match(got[1], want[1], "castBOOLEAN(btrim)")
    for bad in ("yes", "", "t", "10", "truee", "fals", "ｔrue"):
        batch = pa.RecordBatch.from_arrays([pa.array(["1", None, bad], S)], schema=schema)
        with pytest.raises(gandiva.GandivaError, match="ExecutionError: Invalid value for boolean"):
            p.evaluate(batch)
        with pytest.raises(Exception, match="Invalid value for boolean"):
            oracle.project([root], [B], batch)


def test_concurrent_evaluate_from_threads(gandiva, oracle):
    """One Projector and one Filter evaluated from several host threads at once on different
    batches (include/gandiva_b200.h "Threading"): every call gets its own results."""
    import threading
    b = gandiva.TreeExprBuilder()
    schema, outs, _ = cases.case_string_outputs(b)
    exprs = [b.make_expression(r, pa.field("o%d" % i, t)) for i, (r, t) in enumerate(outs)]
    p = gandiva.make_projector(schema, exprs, None)
    cond = cases.q6_condition(b)
    f = gandiva.make_filter(cases.Q6_SCHEMA, b.make_condition(cond))
    work = []
    for k in range(6):
        pb = cases.random_batch(schema, 1500 + 517 * k, seed=100 + k, null_prob=0.1, small=True)
        fb = cases.q6_batch(20_000 + 1111 * k, seed=200 + k, null_permille=10)
        work.append((pb, oracle.project([r for r, _ in outs], [t for _, t in outs], pb),
                     fb, oracle.filter_indices(cond, fb)))
    errors = []

    def run(tid):
        try:
            for it in range(4 if devmem.EMU else 16):
                pb, pwant, fb, fwant = work[(tid + it) % len(work)]
                got = p.evaluate(pb)
                for i, (g, w) in enumerate(zip(got, pwant)):
                    assert_arrays_match(g, w, "thread %d it %d out %d" % (tid, it, i))
                sel = f.evaluate(fb, None)
                assert np.array_equal(sel.to_array().to_numpy().astype(np.uint64), fwant), (tid, it)
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e)[:500])

    threads = [threading.Thread(target=run, args=(t,)) for t in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert errors == [], errors[:3]


def test_cast_string_to_date_and_timestamp(gandiva, oracle):
    """castDATE / castTIMESTAMP(utf8): the accepted spellings against Python's calendar, bit-exact
    with the oracle; malformed or impossible dates raise an ExecutionError in both."""
    import datetime
    rng = np.random.default_rng(12)
    S, D64, TS = pa.string(), pa.date64(), pa.timestamp("ms")
    schema = pa.schema([("s", S)])
    b = gandiva.TreeExprBuilder()
    root_d = b.make_function("castDATE", [cases.F(b, "s", S)], D64)
    root_t = b.make_function("castTIMESTAMP", [cases.F(b, "s", S)], TS)
    p = gandiva.make_projector(schema, [b.make_expression(root_d, pa.field("d", D64)),
                                        b.make_expression(root_t, pa.field("t", TS))], None)
    epoch = datetime.datetime(1970, 1, 1)
    strs, want_d, want_t = [], [], []
    for k in range(3000):
        dt = datetime.datetime(int(rng.integers(1, 9999)), 1, 1) + datetime.timedelta(
            days=int(rng.integers(0, 365)), milliseconds=int(rng.integers(0, 86400000)))
        y, mo, d, hh, mi, ss, ms = dt.year, dt.month, dt.day, dt.hour, dt.minute, dt.second, dt.microsecond // 1000
        style = k % 6
        if style == 0:
            t, tod = "%04d-%02d-%02d" % (y, mo, d), 0
        elif style == 1:
            t, tod = "%d-%d-%d %d:%d" % (y, mo, d, hh, mi), (hh * 60 + mi) * 60000
        elif style == 2:
            t, tod = "%04d-%02d-%02dT%02d:%02d:%02d" % (y, mo, d, hh, mi, ss), ((hh * 60 + mi) * 60 + ss) * 1000
        elif style == 3:
            t, tod = "  %04d-%02d-%02d %02d:%02d:%02d.%03d " % (y, mo, d, hh, mi, ss, ms), ((hh * 60 + mi) * 60 + ss) * 1000 + ms
        elif style == 4:
            t, tod = "%04d-%02d-%02d %02d:%02d:%02d.%d" % (y, mo, d, hh, mi, ss, ms // 100), ((hh * 60 + mi) * 60 + ss) * 1000 + (ms // 100) * 100
        else:
            t, tod = "%04d-%02d-%02d %02d:%02d:%02d.%03d987" % (y, mo, d, hh, mi, ss, ms), ((hh * 60 + mi) * 60 + ss) * 1000 + ms
        day_ms = (datetime.datetime(y, mo, d) - epoch).days * 86400000
        strs.append(None if k % 13 == 5 else t)
        want_d.append(None if k % 13 == 5 else day_ms)
        want_t.append(None if k % 13 == 5 else day_ms + tod)
    batch = pa.RecordBatch.from_arrays([pa.array(strs, S)], schema=schema)
    got = p.evaluate(batch)
    want = oracle.project([root_d, root_t], [D64, TS], batch)
    for g, w in zip(got, want):
        assert_arrays_match(g, w, "cast string to date/timestamp")
    assert got[0].view(pa.int64()).to_pylist() == want_d
    assert got[1].view(pa.int64()).to_pylist() == want_t
    neg = pa.RecordBatch.from_arrays([pa.array(["-0044-03-15", "0000-02-29", "1900-03-01"], S)], schema=schema)
    assert_arrays_match(p.evaluate(neg)[0], oracle.project([root_d], [D64], neg)[0], "BC dates")
    for bad in ("2024-02-30", "2023-02-29", "2024-13-01", "2024-00-10", "2024-01-00", "2024-1-1x", "20240101", "",
                "2024-01-01 24:00", "2024-01-01 12:60", "2024-01-01 12:00:60", "2024-01-01 12", "2024-01-01 12:00:00.",
                "2024-01-01  12:00", "2024/01/01", "1-2", "+2024-01-01"):
        one = pa.RecordBatch.from_arrays([pa.array(["2024-01-01", None, bad], S)], schema=schema)
        with pytest.raises(gandiva.GandivaError, match="ExecutionError: Failed to cast the string to a date"):
            p.evaluate(one)
        with pytest.raises(Exception, match="Failed to cast the string to a date"):
            oracle.project([root_t], [TS], one)


def test_split_part_index_must_be_positive(gandiva, oracle):
    b = gandiva.TreeExprBuilder()
    S, I = pa.string(), pa.int32()
    schema = pa.schema([("s", S), ("k", I)])
    root = b.make_function("split_part", [cases.F(b, "s", S), b.make_literal("-", S), cases.F(b, "k", I)], S)
    p = gandiva.make_projector(schema, [b.make_expression(root, pa.field("o", S))], None)
    ok = pa.RecordBatch.from_arrays([pa.array(["a-b-c", "x", None, "p-q"], S), pa.array([2, 1, 0, 5], I)], schema=schema)
    got, = p.evaluate(ok)   # the row with k = 0 has a NULL string: not evaluated, no error
    assert got.to_pylist() == ["b", "x", None, ""]
    bad = pa.RecordBatch.from_arrays([pa.array(["a-b-c", "x"], S), pa.array([2, 0], I)], schema=schema)
    with pytest.raises(gandiva.GandivaError, match="ExecutionError: Index in split_part must be positive"):
        p.evaluate(bad)
    with pytest.raises(Exception, match="split_part"):
        oracle.project([root], [S], bad)


@pytest.mark.parametrize("n,offset", [(1, 0), (45, 0), (700, 3)])
def test_digests(n, offset, gandiva, oracle):
    """hashSHA256 / hashSHA1 / hashMD5 in the kernels against the oracle (itself pinned to hashlib
    and the published known answers): block-boundary message lengths, multi-block messages, nulls."""
    import hashlib
    b = gandiva.TreeExprBuilder()
    schema, outs, _ = cases.case_digests(b)
    batch = cases.digest_batch(n, seed=n, offset=offset)
    exprs = [b.make_expression(r, pa.field("o%d" % i, t)) for i, (r, t) in enumerate(outs)]
    p = gandiva.make_projector(schema, exprs, None)
    got = p.evaluate(batch)
    want = oracle.project([r for r, _ in outs], [t for _, t in outs], batch)
    for i, (g, w) in enumerate(zip(got, want)):
        assert_arrays_match(g, w, "digests n=%d out=%d" % (n, i))
    s0 = batch.column(0)[0].as_py()
    if s0 is not None:
        assert got[0][0].as_py() == hashlib.sha256(s0.encode()).hexdigest()


def test_date_arithmetic_month_ends(gandiva, oracle):
    """Calendar arithmetic on month-end days (Jan 31 + 1 month, leap days, diffs across them)."""
    b = gandiva.TreeExprBuilder()
    schema, outs, _ = cases.case_date_arith(b)
    exprs = [b.make_expression(r, pa.field("o%d" % i, t)) for i, (r, t) in enumerate(outs)]
    p = gandiva.make_projector(schema, exprs, None)
    for n, seed in ((40, 1), (5003, 2)):
        batch = cases.date_arith_batch(n, seed)
        got = p.evaluate(batch)
        want = oracle.project([r for r, _ in outs], [t for _, t in outs], batch)
        for i, (g, w) in enumerate(zip(got, want)):
            assert_arrays_match(g, w, "date arithmetic n=%d out=%d" % (n, i))


def test_decimal_text_round_trip(gandiva, oracle):
    """castVARCHAR(decimal, n) and castDECIMAL(utf8): kernels against the oracle and both against
    Python's `decimal`; text -> decimal -> text is the identity at the same scale; malformed numbers
    raise an ExecutionError."""
    D = decimal = __import__("decimal").Decimal
    t1, t2, S, L = pa.decimal128(15, 4), pa.decimal128(38, 10), pa.string(), pa.int64()
    schema = pa.schema([("x", t1), ("y", t2), ("s", S)])
    b = gandiva.TreeExprBuilder()
    x, y, s = cases.F(b, "x", t1), cases.F(b, "y", t2), cases.F(b, "s", S)
    fn = b.make_function
    n = lambda v: b.make_literal(v, L)
    outs = [(fn("castVARCHAR", [x, n(64)], S), S), (fn("castVARCHAR", [y, n(64)], S), S), (fn("castVARCHAR", [x, n(5)], S), S),
            (fn("castDECIMAL", [s], pa.decimal128(20, 3)), pa.decimal128(20, 3)),
            (fn("castDECIMAL", [s], pa.decimal128(38, 0)), pa.decimal128(38, 0)),
            (fn("castDECIMAL", [fn("castVARCHAR", [y, n(64)], S)], t2), t2)]
    exprs = [b.make_expression(r, pa.field("o%d" % i, t)) for i, (r, t) in enumerate(outs)]
    p = gandiva.make_projector(schema, exprs, None)
    rng = np.random.default_rng(4)
    base = cases.random_batch(pa.schema([("x", t1), ("y", t2)]), 2000, seed=9, null_prob=0.1)
    texts = []
    for k in range(2000):
        ip = "".join(str(int(d)) for d in rng.integers(0, 10, int(rng.integers(0, 12))))
        fp = "".join(str(int(d)) for d in rng.integers(0, 10, int(rng.integers(0, 8))))
        t = ("-" if k % 3 == 0 else ("+" if k % 7 == 0 else "")) + ip + ("." + fp if (fp or k % 5 == 0) else "")
        if not (ip or fp):
            t = "0"
        texts.append(None if k % 19 == 3 else (" " * (k % 2)) + t + (" " * (k % 3)))
    batch = pa.RecordBatch.from_arrays([base.column(0), base.column(1), pa.array(texts, S)], schema=schema)
    got = p.evaluate(batch)
    want = oracle.project([r for r, _ in outs], [t for _, t in outs], batch)
    for i, (g, w) in enumerate(zip(got, want)):
        assert_arrays_match(g, w, "decimal text out %d" % i)
    xs, ys = batch.column(0).to_pylist(), batch.column(1).to_pylist()
    g0, g1, g3, g5 = got[0].to_pylist(), got[1].to_pylist(), got[3].to_pylist(), got[5].to_pylist()
    for r in range(2000):
        if xs[r] is not None:
            assert D(g0[r]) == xs[r] and g0[r] == format(xs[r], "f")
        if ys[r] is not None:
            assert g1[r] == format(ys[r], "f") and g5[r] == ys[r]
        if texts[r] is not None:
            q = D(texts[r].strip()).quantize(D("0.001"), rounding="ROUND_HALF_UP")
            assert g3[r] == q, (r, texts[r], g3[r], q)
    for bad in ("", " ", "-", "1.2.3", "12a", "1e5", ".", "--1", "1 2"):
        one = pa.RecordBatch.from_arrays([base.column(0).slice(0, 2), base.column(1).slice(0, 2), pa.array(["1.5", bad], S)],
                                         schema=schema)
        with pytest.raises(gandiva.GandivaError, match="ExecutionError: Failed to cast the string to a decimal"):
            p.evaluate(one)
        with pytest.raises(Exception, match="decimal"):
            oracle.project([outs[3][0]], [outs[3][1]], one)
