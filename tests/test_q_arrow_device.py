"""Ingestion through the Arrow C Data / C Device Data interfaces (include/gandiva_b200_arrow.h,
SURVEY.md §8(f)4): struct-typed ArrowDeviceArray in, ArrowDeviceArray out, buffers used in place.
Host arrays go through pyarrow's own exporter / importer; device arrays are assembled by hand
(pyarrow in this image has no CUDA support) from device buffers."""
import ctypes as C
import decimal

import numpy as np
import pyarrow as pa
import pytest

import cases
import devmem
from helpers import assert_arrays_match


def _export_schema(schema: pa.Schema, gandiva):
    c = gandiva.ArrowSchemaC()
    schema._export_to_c(C.addressof(c))
    return c


ALL_TYPES = pa.schema([
    ("b", pa.bool_()), ("i8", pa.int8()), ("u8", pa.uint8()), ("i16", pa.int16()), ("u16", pa.uint16()),
    ("i32", pa.int32()), ("u32", pa.uint32()), ("i64", pa.int64()), ("u64", pa.uint64()),
    ("f", pa.float32()), ("d", pa.float64()), ("s", pa.string()), ("z", pa.binary()),
    ("d32", pa.date32()), ("d64", pa.date64()), ("ts", pa.timestamp("ms")), ("tsz", pa.timestamp("us", tz="UTC")),
    ("t32", pa.time32("ms")), ("t64", pa.time64("us")), ("dec", pa.decimal128(15, 2))])


def test_schema_from_arrow_and_output_schema(gandiva):
    """Format strings of every supported type parse; a projector made on the imported schema
    reports its outputs as an ArrowSchema that pyarrow imports back."""
    lib = gandiva.lib
    c = _export_schema(ALL_TYPES, gandiva)
    h = C.c_void_p()
    gandiva._check(lib.gdv_schema_from_arrow(C.addressof(c), C.byref(h)))
    gandiva._RELEASE_SCHEMA(c.release)(C.byref(c))
    b = gandiva.TreeExprBuilder()
    outs = [(b.make_function("add", [cases.F(b, "i32", pa.int32()), cases.F(b, "i32", pa.int32())], pa.int32()), pa.field("twice", pa.int32())),
            (b.make_function("upper", [cases.F(b, "s", pa.string())], pa.string()), pa.field("up", pa.string())),
            (b.make_function("less_than", [cases.F(b, "dec", pa.decimal128(15, 2)), cases.F(b, "dec", pa.decimal128(15, 2))], pa.bool_()), pa.field("lt", pa.bool_())),
            (cases.F(b, "ts", pa.timestamp("ms")), pa.field("t", pa.timestamp("ms"))),
            (cases.F(b, "dec", pa.decimal128(15, 2)), pa.field("dd", pa.decimal128(15, 2)))]
    exprs = [b.make_expression(r, f) for r, f in outs]
    arr = (C.c_void_p * len(exprs))(*[e._h for e in exprs])
    ph = C.c_void_p()
    cfg = gandiva.Configuration()._c()
    gandiva._check(lib.gdv_projector_make(h, arr, len(exprs), 0, C.byref(cfg), C.byref(ph)))
    out = gandiva.ArrowSchemaC()
    gandiva._check(lib.gdv_projector_output_schema_arrow(ph, C.addressof(out)))
    got = pa.Schema._import_from_c(C.addressof(out))
    assert got == pa.schema([f for _, f in outs])
    lib.gdv_projector_release(ph)
    lib.gdv_schema_release(h)


def test_schema_from_arrow_rejects_unsupported(gandiva):
    lib = gandiva.lib
    for bad in (pa.schema([("l", pa.list_(pa.int32()))]), pa.schema([("x", pa.large_string())]),
                pa.schema([("d", pa.decimal256(40, 2))])):
        c = _export_schema(bad, gandiva)
        h = C.c_void_p()
        assert lib.gdv_schema_from_arrow(C.addressof(c), C.byref(h)) == gandiva.GDV_NOT_IMPLEMENTED
        gandiva._RELEASE_SCHEMA(c.release)(C.byref(c))
    c = gandiva.ArrowSchemaC()
    pa.int32()._export_to_c(C.addressof(c))   # not a struct
    h = C.c_void_p()
    assert lib.gdv_schema_from_arrow(C.addressof(c), C.byref(h)) == gandiva.GDV_INVALID
    gandiva._RELEASE_SCHEMA(c.release)(C.byref(c))


def test_import_validation(gandiva):
    """Import needs no device: mismatching column counts are rejected and leave the producer's
    struct alive; a good import moves it."""
    batch = pa.RecordBatch.from_arrays([pa.array([1, 2, 3], pa.int32())], names=["a"])
    arr, sch = gandiva.ArrowDeviceArrayC(), gandiva.ArrowSchemaC()
    batch._export_to_c_device(C.addressof(arr), C.addressof(sch))
    gandiva._RELEASE_SCHEMA(sch.release)(C.byref(sch))
    with pytest.raises(pa.ArrowInvalid, match="children"):
        gandiva.ArrowDeviceBatch(C.addressof(arr), pa.schema([("a", pa.int32()), ("b", pa.int32())]))
    assert arr.array.release   # still the producer's
    imported = gandiva.ArrowDeviceBatch(C.addressof(arr), batch.schema)
    assert not arr.array.release  # moved
    assert imported.num_rows == 3 and imported.mem_space == 0
    imported.release()
    with pytest.raises(pa.ArrowInvalid, match="released"):
        gandiva.ArrowDeviceBatch(C.addressof(arr), batch.schema)


@pytest.mark.gpu
@pytest.mark.parametrize("n,offset", [(1, 0), (1000, 0), (20011, 7)])
def test_host_round_trip_through_pyarrow(n, offset, gandiva, oracle):
    """pyarrow exporter -> engine -> pyarrow importer, fixed-width + utf8 + bool outputs, sliced."""
    b = gandiva.TreeExprBuilder()
    schema, outs, _ = cases.case_string_outputs(b)
    batch = cases.random_batch(schema, n, seed=n, null_prob=0.15, offset=offset, small=True)
    exprs = [b.make_expression(r, pa.field("o%d" % i, t)) for i, (r, t) in enumerate(outs)]
    p = gandiva.make_projector(schema, exprs, None)
    imported = gandiva.ArrowDeviceBatch.from_record_batch(batch)
    res = p.evaluate_arrow(imported)
    assert res.array.device_type == gandiva.ARROW_DEVICE_CPU and not res.array.sync_event
    got = res.to_record_batch()
    assert got.schema.names == ["o%d" % i for i in range(len(outs))]
    want = oracle.project([r for r, _ in outs], [t for _, t in outs], batch)
    for i, w in enumerate(want):
        assert_arrays_match(got.column(i), w, "arrow host round trip out=%d" % i)
        got.column(i).validate(full=True)
    imported.release()


@pytest.mark.gpu
def test_host_filter_indices_array(gandiva, oracle):
    b = gandiva.TreeExprBuilder()
    batch = cases.q6_batch(50_003, seed=3, null_permille=10)
    cond = cases.q6_condition(b)
    f = gandiva.make_filter(cases.Q6_SCHEMA, b.make_condition(cond))
    imported = gandiva.ArrowDeviceBatch.from_record_batch(batch)
    want = oracle.filter_indices(cond, batch, threads=2)
    for mode, t in (("UINT32", pa.uint32()), ("UINT64", pa.uint64())):
        res = f.evaluate_arrow(imported, mode)
        ts = gandiva.ArrowSchemaC()
        t._export_to_c(C.addressof(ts))
        got = pa.Array._import_from_c_device(res.address, C.addressof(ts))
        assert got.type == t
        assert np.array_equal(got.to_numpy().astype(np.uint64), want)


class _HandMadeDeviceBatch:
    """struct ArrowDeviceArray over device buffers, as a GPU producer would export it."""

    def __init__(self, gandiva, columns, n, offset=0, event_ptr=None):
        self.keep = []
        self.released = []
        nchild = len(columns)
        self.children = (gandiva.ArrowArrayC * nchild)()
        self.child_ptrs = (C.POINTER(gandiva.ArrowArrayC) * nchild)()
        self._child_release = gandiva._RELEASE_ARRAY(lambda a: setattr(a.contents, "release", None))
        for i, (bufs, null_count, length) in enumerate(columns):
            barr = (C.c_void_p * len(bufs))(*[b if b else None for b in bufs])
            self.keep.append(barr)
            c = self.children[i]
            c.length, c.null_count, c.offset = length, null_count, 0
            c.n_buffers, c.buffers = len(bufs), barr
            c.release = C.cast(self._child_release, C.c_void_p)
            self.child_ptrs[i] = C.pointer(c)
        self.array = gandiva.ArrowDeviceArrayC()
        a = self.array.array
        a.length, a.null_count, a.offset = n, 0, offset
        self.parent_bufs = (C.c_void_p * 1)(None)
        a.n_buffers, a.buffers = 1, self.parent_bufs
        a.n_children, a.children = nchild, self.child_ptrs

        def _rel(ptr):
            self.released.append(True)
            ptr.contents.release = None
        self._release = gandiva._RELEASE_ARRAY(_rel)
        a.release = C.cast(self._release, C.c_void_p)
        self.array.device_id = 0
        self.array.device_type = gandiva.ARROW_DEVICE_CUDA
        self.array.sync_event = event_ptr


@pytest.mark.gpu
def test_device_arrays_stay_in_place(gandiva, oracle):
    """ARROW_DEVICE_CUDA in -> ARROW_DEVICE_CUDA out: inputs are read where the producer put
    them (struct-level offset honoured), outputs are device buffers of the engine's pool, the
    producer's release callback runs when the imported batch is released."""
    n_total, off = 100_003, 64
    n = n_total - off
    a = devmem.DevBuf(n_total, np.int32)
    bb = devmem.DevBuf(n_total, np.int32)
    av = devmem.DevBuf((n_total + 31) // 32, np.int32)
    st = devmem.stream()
    gandiva.generate_lineitem(0, 9, 42, 0, n_total, a.ptr, av.ptr, 100, st)
    gandiva.generate_lineitem(0, 10, 42, 0, n_total, bb.ptr, 0, 0, st)
    devmem.synchronize()
    event_ptr = None
    if not devmem.EMU:
        import torch
        ev = torch.cuda.Event()
        ev.record()
        holder = C.c_void_p(ev.cuda_event)
        event_ptr = C.addressof(holder)
    t = pa.int32()
    schema = pa.schema([("a", t), ("b", t)])
    made = _HandMadeDeviceBatch(gandiva, [((av.ptr, a.ptr), -1, n_total), ((0, bb.ptr), 0, n_total)], n, off, event_ptr)
    imported = gandiva.ArrowDeviceBatch(C.addressof(made.array), schema)
    assert imported.mem_space == 1 and imported.num_rows == n
    bld = gandiva.TreeExprBuilder()
    root = bld.make_function("add", [cases.F(bld, "a", t), cases.F(bld, "b", t)], t)
    cond = bld.make_function("less_than", [cases.F(bld, "a", t), cases.F(bld, "b", t)], pa.bool_())
    p = gandiva.make_projector(schema, [bld.make_expression(root, pa.field("c", t))], None)
    f = gandiva.make_filter(schema, bld.make_condition(cond))
    res = p.evaluate_arrow(imported)
    assert res.array.device_type == gandiva.ARROW_DEVICE_CUDA and res.array.array.n_children == 1
    child = res.array.array.children[0].contents
    assert child.length == n and child.n_buffers == 2
    vals = np.empty(n, np.int32)
    vld = np.empty((n + 31) // 32, np.uint32)
    gandiva.memcpy_dtoh(0, vals, child.buffers[1])
    gandiva.memcpy_dtoh(0, vld, child.buffers[0])
    ca, cav = oracle.generate_lineitem(9, 42, 0, n_total, 100, threads=2)
    cb, _ = oracle.generate_lineitem(10, 42, 0, n_total, 0, threads=2)
    batch = pa.RecordBatch.from_arrays(
        [pa.Array.from_buffers(t, n_total, [pa.py_buffer(cav), pa.py_buffer(ca)]),
         pa.Array.from_buffers(t, n_total, [None, pa.py_buffer(cb)])], schema=schema).slice(off)
    want, = oracle.project([root], [t], batch, threads=2)
    got = pa.Array.from_buffers(t, n, [pa.py_buffer(vld), pa.py_buffer(vals)])
    assert_arrays_match(got, want, "device arrow add")
    fres = f.evaluate_arrow(imported, "UINT32")
    k = fres.array.array.length
    idx = np.empty(k, np.uint32)
    gandiva.memcpy_dtoh(0, idx, fres.array.array.buffers[1])
    assert np.array_equal(idx.astype(np.uint64), oracle.filter_indices(cond, batch, threads=2))
    res.release()
    fres.release()
    assert made.released == []
    imported.release()
    assert made.released == [True]
