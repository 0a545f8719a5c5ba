"""The reference's C++ API (namespace gandiva, libgandiva.so) with the RecordBatch resident in HBM:
tests/cpp/device_api_shim.cc copies a host batch to the device (gandiva::CopyToDevice -> arrow buffers
with is_cpu() == false from gandiva::DeviceMemoryManager), calls Filter::Evaluate with a device
SelectionVector and Projector::Evaluate(batch, selection, pool, &outputs), checks that every result
buffer is in device memory, and copies the results back.  Here they are compared with the oracle
(and the same program on the host path must agree too).  P/includes/libgandiva.pxd:218-226,246-248."""
import ctypes as C
import os

import numpy as np
import pyarrow as pa
import pytest

import cases
from helpers import assert_arrays_match

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "tests", "cpp", "_bin", "libdevice_api_shim.so")

SCHEMA = pa.schema([("k", pa.int32()), ("d", pa.float64()), ("q", pa.int64()), ("e", pa.float64()),
                    ("x", pa.decimal128(15, 2))])


def _batch(n, seed, nulls=True):
    import oracle
    p = 20 if nulls else 0
    cols = []
    for kind, f in zip((0, 1, 3, 7, 4), SCHEMA):
        vals, vld = oracle.generate_lineitem(kind, seed, 0, n, p, threads=4)
        cols.append(pa.Array.from_buffers(f.type, n, [pa.py_buffer(vld) if vld is not None else None, pa.py_buffer(vals)]))
    return pa.RecordBatch.from_arrays(cols, schema=SCHEMA)


def _trees(b):
    f = {x.name: b.make_field(x) for x in SCHEMA}
    B, F64, I64, I32 = pa.bool_(), pa.float64(), pa.int64(), pa.int32()
    fn = b.make_function
    cond = b.make_and([fn("greater_than_or_equal_to", [f["k"], b.make_literal(8766, I32)], B),
                       fn("less_than", [f["k"], b.make_literal(9131, I32)], B),
                       fn("greater_than_or_equal_to", [f["d"], b.make_literal(0.05, F64)], B),
                       fn("less_than_or_equal_to", [f["d"], b.make_literal(0.07, F64)], B),
                       fn("less_than", [f["q"], b.make_literal(24, I64)], B)])
    D31 = pa.decimal128(31, 4)
    outs = [(fn("multiply", [f["e"], fn("subtract", [b.make_literal(1.0, F64), f["d"]], F64)], F64), F64),
            (fn("add", [f["q"], f["q"]], I64), I64),
            (b.make_if(fn("greater_than", [f["d"], b.make_literal(0.05, F64)], B), f["e"], b.make_literal(0.0, F64), F64), F64),
            (fn("multiply", [f["x"], f["x"]], D31), D31),
            (b.make_if(fn("less_than_or_equal_to", [f["k"], b.make_literal(10471, I32)], B), f["q"], b.make_literal(None, I64), I64), I64)]
    return cond, outs


def _run_shim(batch, on_device):
    from pyarrow.cffi import ffi
    lib = C.CDLL(SHIM)
    lib.shim_filter_then_project.restype = C.c_int
    lib.shim_filter_then_project.argtypes = [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 5 + [C.c_int]
    structs = [ffi.new("struct ArrowArray*") if i % 2 == 0 else ffi.new("struct ArrowSchema*") for i in range(6)]
    ptr = [int(ffi.cast("uintptr_t", s)) for s in structs]
    batch._export_to_c(ptr[0], ptr[1])
    err = C.create_string_buffer(1024)
    rc = lib.shim_filter_then_project(ptr[0], ptr[1], 1 if on_device else 0, ptr[2], ptr[3], ptr[4], ptr[5], err, 1024)
    assert rc == 0, err.value.decode()
    sel = pa.Array._import_from_c(ptr[2], ptr[3])
    out = pa.RecordBatch._import_from_c(ptr[4], ptr[5])
    return sel, out


def test_shim_is_built():
    """build() compiles the C++ caller against include/gandiva/*.h + libgandiva.so (no GPU needed)."""
    assert os.path.exists(SHIM), "run python gandiva_b200/build.py"
    C.CDLL(SHIM).shim_filter_then_project


@pytest.mark.gpu
@pytest.mark.parametrize("n,nulls", [(1, True), (1000, False), (70_003, True)])
def test_cpp_api_with_device_resident_batch(n, nulls, gandiva, oracle):
    batch = _batch(n, seed=n, nulls=nulls)
    b = gandiva.TreeExprBuilder()
    cond, outs = _trees(b)
    want_idx = oracle.filter_indices(cond, batch, threads=4)
    want = oracle.project([r for r, _ in outs], [t for _, t in outs], batch, selection=want_idx.astype(np.int64), threads=4)
    for on_device in (True, False):
        sel, out = _run_shim(batch, on_device)
        assert sel.type == pa.uint32()
        assert np.array_equal(sel.to_numpy().astype(np.uint64), want_idx), "selection, device=%s" % on_device
        assert out.num_rows == len(want_idx)
        for i in range(len(outs)):
            assert_arrays_match(out.column(i), want[i], "output %d, device=%s" % (i, on_device))
