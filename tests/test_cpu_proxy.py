"""The fused C++ CPU arm of bench.py (oracle/cpu_proxy.cc, kind "fused-cxx-proxy") must compute what
the scalar oracle computes: each hand-fused loop against oracle.project / oracle.filter_indices on
seeded batches whose sizes straddle the 64-row words and the thread partition."""
import numpy as np
import pyarrow as pa
import pytest

import cases
import oracle
from helpers import assert_arrays_match
from oracle import cpu_proxy as px
from oracle.tree import TreeBuilder

SIZES = [0, 1, 63, 64, 65, 1000, 4097, 100_003]


def _bitmap(arr_or_none, n):
    return None if arr_or_none is None else np.frombuffer(arr_or_none, dtype=np.uint8)


@pytest.mark.parametrize("nullp", [0, 10])
@pytest.mark.parametrize("n", SIZES)
def test_q6_filter(n, nullp):
    batch = cases.q6_batch(n, seed=7, null_permille=nullp)
    cols = [batch.column(i).buffers() for i in range(3)]
    vals = [np.frombuffer(c[1], dtype=d) if n else np.zeros(0, d) for c, d in zip(cols, (np.int32, np.float64, np.float64))]
    vl = [_bitmap(c[0], n) for c in cols]
    out = np.zeros(max(n, 1), dtype=np.uint32)
    bits = np.zeros((n + 63) // 64 + 1, dtype=np.uint64)
    cnt = px.q6_filter(vals[0], vals[1], vals[2], vl[0], vl[1], vl[2], n, out, bits)
    want = oracle.filter_indices(cases.q6_condition(TreeBuilder()), batch, threads=2)
    assert np.array_equal(out[:cnt].astype(np.uint64), want)


def test_tree_builder_matches_product_builder(gandiva):
    """The product-free builder serialises to the same s-expression as the product's nodes."""
    for fn, schema in ((cases.q6_condition, cases.Q6_SCHEMA), (cases.comment_condition, cases.COMMENT_SCHEMA)):
        assert oracle.sexpr(fn(TreeBuilder()), schema) == oracle.sexpr(fn(gandiva.TreeExprBuilder()), schema)
    a = [oracle.sexpr(r, cases.Q1_SCHEMA) for r, _ in cases.q1_outputs(TreeBuilder())]
    b = [oracle.sexpr(r, cases.Q1_SCHEMA) for r, _ in cases.q1_outputs(gandiva.TreeExprBuilder())]
    assert a == b


@pytest.mark.parametrize("n", SIZES)
def test_add_i32(n):
    b = TreeBuilder()
    schema, outs, _ = cases.case_arith("add", pa.int32())(b)
    batch = cases.random_batch(schema, n, seed=3, null_prob=0.1)
    bufs = [batch.column(i).buffers() for i in range(2)]
    a, c = [np.frombuffer(x[1], dtype=np.int32)[:n] if n else np.zeros(0, np.int32) for x in bufs]
    out = np.zeros(max(n, 1), dtype=np.int32)
    vout = np.zeros((n + 7) // 8 + 8, dtype=np.uint8)
    px.add_i32(a, c, _bitmap(bufs[0][0], n), _bitmap(bufs[1][0], n), n, out, vout)
    got = pa.Array.from_buffers(pa.int32(), n, [pa.py_buffer(vout), pa.py_buffer(out)])
    want, = oracle.project([outs[0][0]], [pa.int32()], batch, threads=2)
    assert_arrays_match(got, want, "add")


@pytest.mark.parametrize("n", [0, 1, 64, 1000, 50_001])
def test_comment_filter(n):
    batch = cases.comment_batch(n, seed=11) if n else pa.RecordBatch.from_arrays([pa.array([], pa.string())], schema=cases.COMMENT_SCHEMA)
    arr = batch.column(0)
    bufs = arr.buffers()
    offs = np.frombuffer(bufs[1], dtype=np.int32) if bufs[1] is not None and bufs[1].size else np.zeros(1, np.int32)
    data = np.frombuffer(bufs[2], dtype=np.uint8) if bufs[2] is not None and bufs[2].size else np.zeros(1, np.uint8)
    out = np.zeros(max(n, 1), dtype=np.uint32)
    bits = np.zeros((n + 63) // 64 + 1, dtype=np.uint64)
    cnt = px.comment_filter(offs, data, _bitmap(bufs[0], n), n, out, bits)
    want = oracle.filter_indices(cases.comment_condition(TreeBuilder()), batch, threads=2)
    assert np.array_equal(out[:cnt].astype(np.uint64), want)
    if n >= 1000:
        assert cnt > 0


def test_comment_filter_non_ascii_and_long_rows():
    rows = ["special requests", "SPECIAL ééééééééééééééééééé REQUESTS", "é" * 30 + "special requests",
            "x" * 20 + "special requests!", "x" * 17 + "special requests", "requests special", None,
            "ünï special ünï requests ünï", "speCIAL" + "日" * 17 + "reQUESTs", "speCIAL" + "日" * 18 + "reQUESTs", ""]
    arr = pa.array(rows * 9, type=pa.string())
    batch = pa.RecordBatch.from_arrays([arr], schema=cases.COMMENT_SCHEMA)
    n = len(arr)
    bufs = arr.buffers()
    out = np.zeros(n, dtype=np.uint32)
    bits = np.zeros((n + 63) // 64 + 1, dtype=np.uint64)
    cnt = px.comment_filter(np.frombuffer(bufs[1], dtype=np.int32), np.frombuffer(bufs[2], dtype=np.uint8),
                            _bitmap(bufs[0], n), n, out, bits)
    want = oracle.filter_indices(cases.comment_condition(TreeBuilder()), batch, threads=1)
    assert np.array_equal(out[:cnt].astype(np.uint64), want)


@pytest.mark.parametrize("n", [1, 64, 1000, 20_011])
def test_q1_project(n):
    batch = cases.q1_batch(n, seed=5, null_permille=20)
    ins, vins = [], []
    for i in range(8):
        bufs = batch.column(i).buffers()
        ins.append(np.frombuffer(bufs[1], dtype=np.uint8))
        vins.append(_bitmap(bufs[0], n))
    outs_t = cases.q1_outputs(TreeBuilder())
    outs = [np.zeros(n * (t.bit_width // 8), dtype=np.uint8) for _, t in outs_t]
    vouts = [np.zeros((n + 7) // 8 + 8, dtype=np.uint8) for _ in outs_t]
    px.q1_project(ins, vins, n, outs, vouts)
    want = oracle.project([r for r, _ in outs_t], [t for _, t in outs_t], batch, threads=2)
    for k, (_, t) in enumerate(outs_t):
        got = pa.Array.from_buffers(t, n, [pa.py_buffer(vouts[k]), pa.py_buffer(outs[k])])
        assert_arrays_match(got, want[k], "q1 output %d" % k)


def test_generate_matches_oracle_generator():
    n = 10_007
    for kind in (0, 1, 2, 3, 4, 7, 9):
        vals, vld = px.generate(kind, 42, 5, n, 10)
        w, wv = oracle.generate_lineitem(kind, 42, 5, n, 10, threads=1)
        assert np.array_equal(vals.array.reshape(w.shape), w)
        assert np.array_equal(vld.array[: (n + 7) // 8], wv[: (n + 7) // 8])


def test_product_never_loads_the_proxy():
    import os
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gandiva_b200")
    for dp, _, files in os.walk(root):
        if "_build" in dp or "_cubin_cache" in dp:
            continue
        for f in files:
            if f == "build.py":     # the build script compiles the proxy next to the oracle; it is not on any Evaluate path
                continue
            if f.endswith((".py", ".cc", ".h", ".cu", ".cuh")):
                assert "cpu_proxy" not in open(os.path.join(dp, f), errors="replace").read(), f
