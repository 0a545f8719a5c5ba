"""Parity tests of the OPT-IN kernel variants (none is a default): the key-scan string filter
(string_scan bit 4), the two-pass filter (Configuration(loader=3)) and the W-walk filter tiles
(Configuration(stages=W)).  Same bar as test_parity_gpu.py -- bit-exact against the oracle through
the C-ABI -- kept in a file that sorts last so that the default paths are judged first."""
import numpy as np
import pyarrow as pa
import pytest

import cases
import devmem
from helpers import assert_arrays_match

pytestmark = pytest.mark.gpu


KEY_SCAN_CASES = [cases.case_like_scan(p, v, "filter") for p, v in
                  [("%special%requests%", "upper_substr32"), ("%spark%", "plain"), ("spa%ark%fire", "lower"),
                   ("%park%park%", "substr_3_20"), ("%日本語%", "plain"), ("%fire%fox", "btrim"),
                   ("x_y%100%%%park%", "plain"), ("%requests%special%", "upper")]]


@pytest.mark.parametrize("case", KEY_SCAN_CASES, ids=[c.__name__ for c in KEY_SCAN_CASES])
def test_key_scan_filter(case, gandiva, oracle):
    """string_scan bit 4: the filter is driven by the occurrences of a literal LIKE segment in the
    column's bytes (rows without one are never looked at).  Same answers as the oracle on sparse
    and dense matches (list overflow -> second pass), rows longer than a segment, non-ASCII text,
    sliced arrays, 16/32/64-bit indices and a row base."""
    b = gandiva.TreeExprBuilder()
    schema, outs, _ = case(b)
    cond = outs[0][0]
    f = gandiva.make_filter(schema, b.make_condition(cond), gandiva.Configuration(string_scan=16))
    assert "key-scan string Filter" in f.llvm_ir
    for n, seed, offset, dense, long_rows in [(1, 1, 0, False, False), (64, 1, 0, False, False),
                                              (5000, 2, 0, False, False), (20011, 3, 7, True, False),
                                              (9000, 4, 1, False, True), (60_001, 5, 3, True, True)]:
        batch = cases.like_scan_batch(n, seed, offset=offset, dense=dense, long_rows=long_rows)
        want = oracle.filter_indices(cond, batch, threads=4)
        for dtype in ("int32", "int64") if n != 5000 else ("int16", "int32"):
            sel = f.evaluate(batch, None, dtype)
            assert sel.num_slots == len(want), (n, dtype, sel.num_slots, len(want))
            assert np.array_equal(sel.to_array().to_numpy().astype(np.uint64), want), (n, dtype)


def test_key_scan_filter_conjunction_and_fallback(gandiva, oracle):
    """The LIKE may sit anywhere on the AND spine next to other predicates (evaluated for the
    candidate rows only); conditions that do not imply a key keep the row-driven kernel."""
    b = gandiva.TreeExprBuilder()
    cond = cases.comment_condition(b)
    cfg = gandiva.Configuration(string_scan=16)
    f = gandiva.make_filter(cases.COMMENT_SCHEMA, b.make_condition(cond), cfg)
    assert "key-scan string Filter" in f.llvm_ir
    for n in (70_001, 300_000):
        batch = cases.comment_batch(n, seed=n)
        want = oracle.filter_indices(cond, batch, threads=4)
        sel = f.evaluate(batch)
        assert len(want) > 0 and np.array_equal(sel.to_array().to_numpy().astype(np.uint64), want)
    S, I, B = pa.string(), pa.int32(), pa.bool_()
    schema = pa.schema([("s", S), ("k", I)])
    s, k = cases.F(b, "s", S), cases.F(b, "k", I)
    like = b.make_function("like", [s, b.make_literal("%park%", S)], B)
    small = b.make_function("less_than", [k, b.make_literal(0, I)], B)
    both = b.make_and([small, b.make_and([like, b.make_function("isnotnull", [k], B)])])
    either = b.make_or([small, like])
    batch = cases.random_batch(schema, 30_011, seed=9, null_prob=0.1, offset=5)
    f_and = gandiva.make_filter(schema, b.make_condition(both), cfg)
    assert "key-scan string Filter" in f_and.llvm_ir
    assert np.array_equal(f_and.evaluate(batch).to_array().to_numpy().astype(np.uint64),
                          oracle.filter_indices(both, batch, threads=4))
    # other ways of saying "the column holds this literal": is_substr / starts_with / ends_with / equal
    for fname, lit, view in (("is_substr", "park", None), ("starts_with", "SPECIAL", "upper"),
                             ("ends_with", "fire", "btrim"), ("equal", "special requests", "lower")):
        arg = s if view is None else b.make_function(view, [s], S)
        c2 = b.make_and([b.make_function(fname, [arg, b.make_literal(lit, S)], B), b.make_function("isnotnull", [k], B)])
        f2 = gandiva.make_filter(schema, b.make_condition(c2), cfg)
        assert "key-scan string Filter" in f2.llvm_ir, fname
        want2 = oracle.filter_indices(c2, batch, threads=4)
        assert np.array_equal(f2.evaluate(batch).to_array().to_numpy().astype(np.uint64), want2), fname
    f_or = gandiva.make_filter(schema, b.make_condition(either), cfg)
    assert "key-scan string Filter" not in f_or.llvm_ir   # an OR does not imply the key
    assert np.array_equal(f_or.evaluate(batch).to_array().to_numpy().astype(np.uint64),
                          oracle.filter_indices(either, batch, threads=4))


def test_key_scan_filter_dense_and_empty(gandiva, oracle):
    """More accepted rows per warp segment than the shared-memory list holds (second, direct-write
    pass over the tile), a column of empty strings (no bytes at all), and a bounded vector."""
    b = gandiva.TreeExprBuilder()
    S, B = pa.string(), pa.bool_()
    schema = pa.schema([("s", S)])
    cond = b.make_function("like", [cases.F(b, "s", S), b.make_literal("%ark%", S)], B)
    f = gandiva.make_filter(schema, b.make_condition(cond), gandiva.Configuration(string_scan=16))
    assert "key-scan string Filter" in f.llvm_ir
    rng = np.random.default_rng(3)
    rows = [None if rng.random() < 0.02 else ("ark" if rng.random() < 0.97 else "xy") for _ in range(40_003)]
    batch = pa.RecordBatch.from_arrays([pa.array(rows, S)], schema=schema)
    want = oracle.filter_indices(cond, batch, threads=4)
    assert len(want) > 35_000
    sel = f.evaluate(batch)
    assert np.array_equal(sel.to_array().to_numpy().astype(np.uint64), want)
    empty = pa.RecordBatch.from_arrays([pa.array([""] * 1000 + [None] * 5, S)], schema=schema)
    assert f.evaluate(empty).num_slots == 0
    # big segments / big CTAs (the variant large batches get): rows_per_thread = KB per warp segment
    for bt, seg_kb in ((1024, 64), (64, 1), (512, 16)):
        f2 = gandiva.make_filter(schema, b.make_condition(cond),
                                 gandiva.Configuration(string_scan=16, block_threads=bt, rows_per_thread=seg_kb))
        assert np.array_equal(f2.evaluate(batch).to_array().to_numpy().astype(np.uint64), want), (bt, seg_kb)
    # "arkark": two occurrences in one row, the row is reported once
    twice = pa.RecordBatch.from_arrays([pa.array(["arkark", "xarkxxark", "ar", "k", "ark"] * 700, S)], schema=schema)
    sel = f.evaluate(twice)
    assert np.array_equal(sel.to_array().to_numpy().astype(np.uint64), oracle.filter_indices(cond, twice))


@pytest.mark.parametrize("nullp", [0, 15])
def test_two_pass_filter(nullp, gandiva, oracle):
    """Configuration(loader=3) on device batches: condition -> truth bitmap with the projector
    kernel, bitmap -> ordered SelectionVector with gdv_bitmap_to_sel.  Same indices as the oracle
    for every index width, with a row base and a bounded vector, at tile and word boundaries."""
    b = gandiva.TreeExprBuilder()
    cond = cases.q6_condition(b)
    f = gandiva.make_filter(cases.Q6_SCHEMA, b.make_condition(cond), gandiva.Configuration(loader=3))
    st = devmem.stream()
    for n in (1, 31, 32, 33, 4095, 131072, 131073, 300_011):
        ship, disc, qty = devmem.DevBuf(n, np.int32), devmem.DevBuf(n, np.float64), devmem.DevBuf(n, np.float64)
        vl = [devmem.DevBuf((n + 31) // 32, np.int32) if nullp else None for _ in range(3)]
        for kind, t, v in ((0, ship, vl[0]), (1, disc, vl[1]), (2, qty, vl[2])):
            gandiva.generate_lineitem(0, kind, 42, 0, n, t.ptr, v.ptr if v is not None else 0, nullp, st)
        cols = [(v.ptr if v is not None else 0, t.ptr, 0, 0) for t, v in zip((ship, disc, qty), vl)]
        batch = cases.q6_batch(n, seed=42, null_permille=nullp)
        want = oracle.filter_indices(cond, batch, threads=4)
        modes = [("UINT32", np.uint32, 0), ("UINT64", np.uint64, 7_000_000_000)]
        if n <= 65536:
            modes.append(("UINT16", np.uint16, 0))
        for mode, npdt, base in modes:
            out = devmem.DevBuf(n + 8, npdt, fill=0)
            cnt = devmem.DevBuf(1, np.int64, fill=0)
            f.evaluate_device(n, cols, out.ptr, n, mode, st, cnt.ptr, index_base=base)
            count = f.sync(st)
            assert count == len(want) == int(cnt.numpy()[0]), (n, mode)
            assert np.array_equal(out.numpy()[:count].astype(np.uint64), want + base), (n, mode)
        if n > 1000:
            cap = max(1, len(want) // 2)
            out = devmem.DevBuf(cap + 16, np.int64, fill=-1)
            cnt = devmem.DevBuf(1, np.int64, fill=0)
            f.evaluate_device(n, cols, out.ptr, cap, "UINT64|BOUNDED", st, cnt.ptr)
            assert f.sync(st) == len(want)
            got = out.numpy()
            assert np.array_equal(got[:cap].astype(np.uint64), want[:cap]) and (got[cap:] == -1).all()
    assert "gdv_project_expr_" in f.kernel_info["name"]


@pytest.mark.parametrize("walk,bt", [(2, 256), (4, 256), (8, 128), (4, 64)])
def test_filter_walk_variant(walk, bt, gandiva, oracle):
    """Configuration(stages = W) on a fixed-width filter: every warp walks W 1024-row chunks per
    tile (same fused kernel, smaller CTAs).  Same indices as the oracle, tails and bounded vectors."""
    b = gandiva.TreeExprBuilder()
    cond = cases.q6_condition(b)
    f = gandiva.make_filter(cases.Q6_SCHEMA, b.make_condition(cond), gandiva.Configuration(stages=walk, block_threads=bt))
    assert "mymask[%d]" % walk in f.llvm_ir
    for n, nullp in ((1, 0), (1023, 10), (1024 * walk * (bt // 32) + 5, 0), (300_007, 15)):
        batch = cases.q6_batch(n, seed=7, null_permille=nullp)
        want = oracle.filter_indices(cond, batch, threads=4)
        for dtype in ("int32", "int64"):
            sel = f.evaluate(batch, None, dtype)
            assert np.array_equal(sel.to_array().to_numpy().astype(np.uint64), want), (n, nullp, dtype)


