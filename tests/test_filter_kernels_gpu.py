"""Parity tests of the Filter kernel shapes: the key-driven string filter (default wherever the
condition implies a literal; Configuration(string_scan=4) keeps the row-driven kernel, which must
agree), the fixed-width filter with W chunks per warp (Configuration(stages=W); W=4 is what big
batches get) and the validity words hoisted out of the row loop.  Same bar as test_parity_gpu.py:
bit-exact against the oracle through the C-ABI."""
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
def test_key_driven_filter(case, gandiva, oracle):
    """The filter is driven by the occurrences of a literal LIKE segment in the column's bytes (rows
    without one are never looked at).  Same answers as the oracle -- and as the row-driven kernel --
    on sparse and dense matches (anchor list drained many times / overflowing), long rows, non-ASCII
    text, sliced arrays and 16/32/64-bit indices."""
    b = gandiva.TreeExprBuilder()
    schema, outs, _ = case(b)
    cond = outs[0][0]
    f = gandiva.make_filter(schema, b.make_condition(cond))
    assert "key-driven string Filter" in f.llvm_ir
    f_rows = gandiva.make_filter(schema, b.make_condition(cond), gandiva.Configuration(string_scan=4))
    assert "key-driven string Filter" not in f_rows.llvm_ir
    for n, seed, offset, dense, long_rows in [(1, 1, 0, False, False), (64, 1, 0, False, False),
                                              (5000, 2, 0, False, False), (20011, 3, 7, True, False),
                                              (9000, 4, 1, False, True), (60_001, 5, 3, True, True)]:
        batch = cases.like_scan_batch(n, seed, offset=offset, dense=dense, long_rows=long_rows)
        want = oracle.filter_indices(cond, batch, threads=4)
        for dtype in ("int32", "int64") if n != 5000 else ("int16", "int32"):
            sel = f.evaluate(batch, None, dtype)
            assert sel.num_slots == len(want), (n, dtype, sel.num_slots, len(want))
            assert np.array_equal(sel.to_array().to_numpy().astype(np.uint64), want), (n, dtype)
        assert np.array_equal(f_rows.evaluate(batch).to_array().to_numpy().astype(np.uint64), want), n


def test_key_driven_filter_conjunction_and_fallback(gandiva, oracle):
    """The LIKE may sit anywhere on the AND spine next to other predicates (evaluated for the
    candidate rows only); conditions that do not imply a key keep the row-driven kernel."""
    b = gandiva.TreeExprBuilder()
    cond = cases.comment_condition(b)
    cfg = None
    f = gandiva.make_filter(cases.COMMENT_SCHEMA, b.make_condition(cond), cfg)
    assert "key-driven string Filter" in f.llvm_ir and "'REQUESTS'" in f.llvm_ir
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
    assert "key-driven string Filter" in f_and.llvm_ir
    assert np.array_equal(f_and.evaluate(batch).to_array().to_numpy().astype(np.uint64),
                          oracle.filter_indices(both, batch, threads=4))
    # other ways of saying "the column holds this literal": is_substr / starts_with / ends_with / equal
    for fname, lit, view in (("is_substr", "park", None), ("starts_with", "SPECIAL", "upper"),
                             ("ends_with", "fire", "btrim"), ("equal", "special requests", "lower")):
        arg = s if view is None else b.make_function(view, [s], S)
        c2 = b.make_and([b.make_function(fname, [arg, b.make_literal(lit, S)], B), b.make_function("isnotnull", [k], B)])
        f2 = gandiva.make_filter(schema, b.make_condition(c2), cfg)
        assert "key-driven string Filter" in f2.llvm_ir, fname
        want2 = oracle.filter_indices(c2, batch, threads=4)
        assert np.array_equal(f2.evaluate(batch).to_array().to_numpy().astype(np.uint64), want2), fname
    f_or = gandiva.make_filter(schema, b.make_condition(either), cfg)
    assert "key-driven string Filter" not in f_or.llvm_ir   # an OR does not imply the key
    assert np.array_equal(f_or.evaluate(batch).to_array().to_numpy().astype(np.uint64),
                          oracle.filter_indices(either, batch, threads=4))


def test_key_driven_filter_dense_and_empty(gandiva, oracle):
    """More anchors per warp than the shared-memory list holds (many drains; with every chunk full
    of matches the warp evaluates all its rows), a column of empty strings (no bytes at all), a
    3-byte key (halfword test + verification) and rows with several occurrences."""
    b = gandiva.TreeExprBuilder()
    S, B = pa.string(), pa.bool_()
    schema = pa.schema([("s", S)])
    cond = b.make_function("like", [cases.F(b, "s", S), b.make_literal("%ark%", S)], B)
    f = gandiva.make_filter(schema, b.make_condition(cond))
    assert "key-driven string Filter" in f.llvm_ir
    rng = np.random.default_rng(3)
    rows = [None if rng.random() < 0.02 else ("ark" if rng.random() < 0.97 else "xy") for _ in range(40_003)]
    batch = pa.RecordBatch.from_arrays([pa.array(rows, S)], schema=schema)
    want = oracle.filter_indices(cond, batch, threads=4)
    assert len(want) > 35_000
    sel = f.evaluate(batch)
    assert np.array_equal(sel.to_array().to_numpy().astype(np.uint64), want)
    empty = pa.RecordBatch.from_arrays([pa.array([""] * 1000 + [None] * 5, S)], schema=schema)
    assert f.evaluate(empty).num_slots == 0
    for bt, w in ((1024, 1), (64, 2), (512, 4), (256, 2)):     # stages = 1024-row chunks per warp and tile
        f2 = gandiva.make_filter(schema, b.make_condition(cond), gandiva.Configuration(block_threads=bt, stages=w))
        assert "u32 mymask[%d];" % w in f2.llvm_ir
        assert np.array_equal(f2.evaluate(batch).to_array().to_numpy().astype(np.uint64), want), (bt, w)
    # a long key in every row: far more than 128 anchors per 2 KB of bytes -> the all-rows fallback
    cond8 = b.make_function("like", [cases.F(b, "s", S), b.make_literal("%requests%", S)], B)
    f8 = gandiva.make_filter(schema, b.make_condition(cond8))
    assert "key-driven string Filter" in f8.llvm_ir
    full = pa.RecordBatch.from_arrays([pa.array(["requestsrequests", "xrequests", None, "request"] * 9000, S)], schema=schema)
    assert np.array_equal(f8.evaluate(full).to_array().to_numpy().astype(np.uint64), oracle.filter_indices(cond8, full, threads=4))
    # "arkark": two occurrences in one row, the row is reported once
    twice = pa.RecordBatch.from_arrays([pa.array(["arkark", "xarkxxark", "ar", "k", "ark"] * 700, S)], schema=schema)
    sel = f.evaluate(twice)
    assert np.array_equal(sel.to_array().to_numpy().astype(np.uint64), oracle.filter_indices(cond, twice))


@pytest.mark.parametrize("walk,bt", [(1, 256), (2, 256), (4, 256), (8, 128), (4, 64), (1, 1024)])
def test_filter_walk_variant(walk, bt, gandiva, oracle):
    """Configuration(stages = W) on a fixed-width filter: every warp walks W 1024-row chunks per
    tile (same fused kernel, smaller CTAs).  Same indices as the oracle, tails and bounded vectors."""
    b = gandiva.TreeExprBuilder()
    cond = cases.q6_condition(b)
    f = gandiva.make_filter(cases.Q6_SCHEMA, b.make_condition(cond), gandiva.Configuration(stages=walk, block_threads=bt))
    assert "u32 mymask[%d];" % walk in f.llvm_ir
    for n, nullp in ((1, 0), (1023, 10), (1024 * walk * (bt // 32) + 5, 0), (300_007, 15)):
        batch = cases.q6_batch(n, seed=7, null_permille=nullp)
        want = oracle.filter_indices(cond, batch, threads=4)
        for dtype in ("int32", "int64"):
            sel = f.evaluate(batch, None, dtype)
            assert np.array_equal(sel.to_array().to_numpy().astype(np.uint64), want), (n, nullp, dtype)


def test_filter_hoisted_validity(gandiva, oracle):
    """Columns whose NULL makes the condition not-true lose their per-row validity loads: their bitmap
    words are ANDed into the keep-mask after the row loop (gdv_ldwin_rows).  Q6 hoists all three
    columns; an OR keeps the per-row form for the column only one side needs; sliced batches (validity
    offsets that are not multiples of 8 or 32) and every tail length must agree with the oracle."""
    b = gandiva.TreeExprBuilder()
    cond = cases.q6_condition(b)
    f = gandiva.make_filter(cases.Q6_SCHEMA, b.make_condition(cond))
    src = f.llvm_ir
    assert src.count("gdv_ldwin_rows(") == 3 and "gdv_ldwin(in_vp" not in src
    full = cases.q6_batch(70_000, seed=11, null_permille=200)
    for off, n in ((0, 70_000), (1, 1), (3, 33), (5, 1024), (7, 1025), (31, 2047), (33, 32_768), (64, 65_536 + 13), (13, 69_000)):
        batch = full.slice(off, n)
        want = oracle.filter_indices(cond, batch, threads=2)
        got = f.evaluate(batch).to_array().to_numpy().astype(np.uint64)
        assert np.array_equal(got, want), (off, n)
    # a OR (b AND c): nothing is strict in every branch except what both sides share
    t, B = pa.int32(), pa.bool_()
    schema = pa.schema([("a", t), ("b", t), ("c", t)])
    A_, B_, C_ = (cases.F(b, x, t) for x in "abc")
    lt = lambda x, v: b.make_function("less_than", [x, b.make_literal(v, t)], B)   # noqa: E731
    mixed = b.make_and([lt(A_, 0), b.make_or([lt(B_, 0), b.make_and([lt(C_, 0), b.make_function("isnull", [B_], B)])])])
    fm = gandiva.make_filter(schema, b.make_condition(mixed))
    assert fm.llvm_ir.count("gdv_ldwin_rows(") == 1      # only `a` is strict
    rb = cases.random_batch(schema, 50_021, seed=4, null_prob=0.3, offset=9)
    assert np.array_equal(fm.evaluate(rb).to_array().to_numpy().astype(np.uint64), oracle.filter_indices(mixed, rb, threads=2))


