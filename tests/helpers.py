"""Comparison helpers for parity tests."""
from __future__ import annotations

import numpy as np
import pyarrow as pa


def validity_np(arr: pa.Array) -> np.ndarray:
    return np.asarray(arr.is_valid().to_numpy(zero_copy_only=False), dtype=bool)


def ulp_diff(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Distance in units in the last place between two float arrays of the same dtype."""
    it = np.int32 if a.dtype == np.float32 else np.int64
    ai = a.view(it).astype(np.int64)
    bi = b.view(it).astype(np.int64)
    sign = np.int64(1) << (31 if a.dtype == np.float32 else 63)
    # map the sign-magnitude float ordering onto a monotonic integer line
    ai = np.where(ai < 0, -(ai & ~sign) if a.dtype == np.float64 else -(ai & 0x7fffffff), ai)
    bi = np.where(bi < 0, -(bi & ~sign) if a.dtype == np.float64 else -(bi & 0x7fffffff), bi)
    return np.abs(ai - bi)


def assert_arrays_match(got: pa.Array, want: pa.Array, what: str = "", float_ulps: int = 0) -> None:
    """Validity must match bit for bit.  Values are compared on valid slots only (Arrow leaves
    null slots undefined): bit-exact for integer / decimal / string / date / bool, and within
    `float_ulps` ULP for float32/float64 (BASELINE.json allows 1; the engine compiles with
    --fmad=false so the tests ask for 0 unless stated)."""
    assert got.type == want.type, "%s: type %s != %s" % (what, got.type, want.type)
    assert len(got) == len(want), "%s: length %d != %d" % (what, len(got), len(want))
    gv, wv = validity_np(got), validity_np(want)
    if not np.array_equal(gv, wv):
        bad = np.nonzero(gv != wv)[0]
        raise AssertionError("%s: validity differs at %d rows, first %s (got %s want %s)" % (
            what, len(bad), bad[:5], gv[bad[:5]], wv[bad[:5]]))
    t = got.type
    if pa.types.is_floating(t):
        g = got.fill_null(0).to_numpy(zero_copy_only=False)
        w = want.fill_null(0).to_numpy(zero_copy_only=False)
        g, w = g[gv], w[gv]
        both_nan = np.isnan(g) & np.isnan(w)
        d = ulp_diff(np.ascontiguousarray(g), np.ascontiguousarray(w))
        bad = np.nonzero((d > float_ulps) & ~both_nan)[0]
        if len(bad):
            raise AssertionError("%s: %d float values differ by more than %d ULP, first: got %r want %r" % (
                what, len(bad), float_ulps, g[bad[:3]], w[bad[:3]]))
        return
    if pa.types.is_temporal(t):
        # compare the stored integers: to_pylist() raises on values outside datetime's range
        it = pa.int32() if t.bit_width == 32 else pa.int64()
        got, want = got.view(it), want.view(it)
    gl = got.to_pylist()
    wl = want.to_pylist()
    if gl != wl:
        for i, (x, y) in enumerate(zip(gl, wl)):
            if x != y:
                raise AssertionError("%s: row %d got %r want %r" % (what, i, x, y))
