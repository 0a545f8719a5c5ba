"""ctypes wrapper over oracle/libgdv_oracle.so (scalar C++ interpreter, gdv_oracle.cc).

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.  The tree is taken from the Python-side
structure of the harness nodes (`kind`, `dtype`, `children`, `payload`) and serialised to an
s-expression, so the oracle shares no code with the product's C++ node classes.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
from typing import Any, Sequence

import numpy as np
import pyarrow as pa

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgdv_oracle.so")


class _Column(C.Structure):
    _fields_ = [("validity", C.c_void_p), ("values", C.c_void_p), ("var_data", C.c_void_p),
                ("offset", C.c_int64), ("var_data_size", C.c_int64)]


def _load() -> C.CDLL:
    if not os.path.exists(_LIB_PATH):
        raise ImportError("oracle: %s missing; run `python -m gandiva_b200.build`" % _LIB_PATH)
    L = C.CDLL(_LIB_PATH)
    L.orc_parse.restype = C.c_void_p
    L.orc_parse.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    L.orc_free.argtypes = [C.c_void_p]
    L.orc_project.restype = C.c_int
    L.orc_project.argtypes = [C.c_void_p, C.POINTER(_Column), C.c_int64, C.c_void_p, C.c_int64,
                              C.c_void_p, C.c_void_p, C.c_int]
    L.orc_project_string.restype = C.c_int
    L.orc_project_string.argtypes = [C.c_void_p, C.POINTER(_Column), C.c_int64, C.c_void_p,
                                     C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.POINTER(C.c_int64)]
    L.orc_filter.restype = C.c_int
    L.orc_filter.argtypes = [C.c_void_p, C.POINTER(_Column), C.c_int64, C.c_void_p,
                             C.POINTER(C.c_int64), C.c_int]
    L.orc_generate_lineitem.argtypes = [C.c_int, C.c_uint64, C.c_int64, C.c_int64, C.c_void_p,
                                        C.c_void_p, C.c_int, C.c_int]
    L.orc_hardware_threads.restype = C.c_int
    return L


_lib = _load()


class OracleExecutionError(Exception):
    def __init__(self, code: int):
        super().__init__({1: "divide by zero error", 4: "Failed to cast the string to an integer", 5: "Failed to cast the string to a date / timestamp", 6: "Index in split_part must be positive", 7: "Failed to cast the string to a decimal", 8: "Failed to cast the string to a float", 9: "Invalid value for boolean", 10: "Output buffer length can't be negative", 11: "Start position must be greater than 0", 12: "Factorial of negative number not exist!", 13: "Factorial of number greater than 20 not supported!", 14: "Error parsing value for given format (to_date)"}.get(code, "oracle error %d" % code))
        self.code = code


def hardware_threads() -> int:
    return int(_lib.orc_hardware_threads())


# ---- tree -> s-expression ----------------------------------------------------------------
def _type_tok(t: pa.DataType) -> str:
    if pa.types.is_decimal128(t):
        return "decimal128:%d:%d" % (t.precision, t.scale)
    m = {pa.bool_(): "bool", pa.int8(): "int8", pa.int16(): "int16", pa.int32(): "int32",
         pa.int64(): "int64", pa.uint8(): "uint8", pa.uint16(): "uint16", pa.uint32(): "uint32",
         pa.uint64(): "uint64", pa.float32(): "float32", pa.float64(): "float64",
         pa.string(): "utf8", pa.binary(): "binary", pa.date32(): "date32", pa.date64(): "date64"}
    if t in m:
        return m[t]
    if pa.types.is_timestamp(t):
        return "timestamp"
    if pa.types.is_time32(t):
        return "time32"
    if pa.types.is_time64(t):
        return "time64"
    raise NotImplementedError("oracle: type %s" % t)


def _lit_tok(t: pa.DataType, v: Any) -> str:
    if v is None:
        return "null"
    if pa.types.is_string(t) or pa.types.is_binary(t):
        raw = v.encode("utf-8") if isinstance(v, str) else bytes(v)
        return "h" + raw.hex()
    if pa.types.is_float32(t):
        return "x%08x" % struct.unpack("<I", struct.pack("<f", float(v)))[0]
    if pa.types.is_float64(t):
        return "x%016x" % struct.unpack("<Q", struct.pack("<d", float(v)))[0]
    if pa.types.is_boolean(t):
        return "1" if v else "0"
    return str(int(v))


def sexpr(node: Any, schema: pa.Schema) -> str:
    k = node.kind
    if k == "field":
        return "(field %d %s)" % (schema.get_field_index(node.payload), _type_tok(node.dtype))
    if k == "literal":
        return "(lit %s %s)" % (_type_tok(node.dtype), _lit_tok(node.dtype, node.payload))
    kids = " ".join(sexpr(c, schema) for c in node.children)
    if k == "function":
        return "(fn %s %s %s)" % (node.payload, _type_tok(node.dtype), kids)
    if k == "if":
        return "(if %s %s)" % (_type_tok(node.dtype), kids)
    if k in ("and", "or"):
        return "(%s %s)" % (k, kids)
    if k == "in":
        vt, vals = node.payload
        return "(in %s %s %s)" % (_type_tok(vt), kids, " ".join(_lit_tok(vt, v) for v in vals))
    raise NotImplementedError(k)


class _Parsed:
    def __init__(self, text: str):
        err = C.create_string_buffer(512)
        self.h = _lib.orc_parse(text.encode(), err, 512)
        if not self.h:
            raise ValueError("oracle parse error: %s in %s" % (err.value.decode(), text))

    def __del__(self):
        try:
            _lib.orc_free(self.h)
        except Exception:
            pass


def _columns(batch: pa.RecordBatch):
    cols = (_Column * max(batch.num_columns, 1))()
    keep = []
    for i in range(batch.num_columns):
        arr = batch.column(i)
        bufs = arr.buffers()
        keep.append(bufs)
        cols[i].validity = bufs[0].address if bufs[0] is not None else None
        cols[i].values = bufs[1].address if len(bufs) > 1 and bufs[1] is not None else None
        if len(bufs) > 2 and bufs[2] is not None:
            cols[i].var_data = bufs[2].address
            cols[i].var_data_size = bufs[2].size
        cols[i].offset = arr.offset
    return cols, keep


def project(roots: Sequence[Any], result_types: Sequence[pa.DataType], batch: pa.RecordBatch,
            selection: np.ndarray | None = None, threads: int = 1) -> list:
    """Evaluate each root node over `batch` (optionally over the rows in `selection`)."""
    cols, keep = _columns(batch)
    sel = None
    n = batch.num_rows
    count = n
    if selection is not None:
        sel = np.ascontiguousarray(selection, dtype=np.int64)
        count = len(sel)
    out = []
    for root, t in zip(roots, result_types):
        parsed = _Parsed(sexpr(root, batch.schema))
        vbuf = np.zeros((count + 7) // 8 + 8, dtype=np.uint8)
        selp = sel.ctypes.data_as(C.c_void_p) if sel is not None else None
        if pa.types.is_string(t) or pa.types.is_binary(t):
            offs = np.zeros(count + 1, dtype=np.int32)
            cap = 1 << 16
            while True:
                data = np.zeros(cap, dtype=np.uint8)
                needed = C.c_int64()
                rc = _lib.orc_project_string(parsed.h, cols, n, selp, count, vbuf.ctypes.data,
                                             offs.ctypes.data, data.ctypes.data, cap,
                                             C.byref(needed))
                if rc == -1:
                    cap = int(needed.value) + 16
                    vbuf[:] = 0
                    continue
                if rc != 0:
                    raise OracleExecutionError(rc)
                break
            out.append(pa.Array.from_buffers(t, count, [pa.py_buffer(vbuf), pa.py_buffer(offs),
                                                        pa.py_buffer(data)]))
            continue
        if pa.types.is_boolean(t):
            dbuf = np.zeros((count + 7) // 8 + 8, dtype=np.uint8)
        else:
            dbuf = np.zeros(max(count * (t.bit_width // 8), 8), dtype=np.uint8)
        rc = _lib.orc_project(parsed.h, cols, n, selp, count, vbuf.ctypes.data, dbuf.ctypes.data,
                              threads)
        if rc != 0:
            raise OracleExecutionError(rc)
        out.append(pa.Array.from_buffers(t, count, [pa.py_buffer(vbuf), pa.py_buffer(dbuf)]))
    return out


def filter_indices(root: Any, batch: pa.RecordBatch, threads: int = 1) -> np.ndarray:
    """Ascending indices (uint64) of the rows where the condition is valid and true."""
    cols, keep = _columns(batch)
    parsed = _Parsed(sexpr(root, batch.schema))
    idx = np.zeros(max(batch.num_rows, 1), dtype=np.uint64)
    cnt = C.c_int64()
    rc = _lib.orc_filter(parsed.h, cols, batch.num_rows, idx.ctypes.data, C.byref(cnt), threads)
    if rc != 0:
        raise OracleExecutionError(rc)
    return idx[: cnt.value].copy()


_LI_DTYPE = {0: np.int32, 1: np.float64, 2: np.float64, 3: np.int64, 7: np.float64, 8: np.float64,
             9: np.int32, 10: np.int32}


def generate_lineitem(kind: int, seed: int, first_row: int, num_rows: int, null_permille: int = 0,
                      threads: int = 1):
    """CPU twin of gdv_generate_lineitem: returns (values ndarray, validity ndarray|None).
    Decimal kinds (4,5,6) return a (num_rows, 2) uint64 array (lo, hi)."""
    if kind in (4, 5, 6):
        vals = np.zeros((num_rows, 2), dtype=np.uint64)
    else:
        vals = np.zeros(num_rows, dtype=_LI_DTYPE.get(kind, np.int32))
    vld = np.zeros((num_rows + 7) // 8 + 8, dtype=np.uint8) if null_permille > 0 else None
    _lib.orc_generate_lineitem(kind, seed, first_row, num_rows, vals.ctypes.data,
                               vld.ctypes.data if vld is not None else None, null_permille, threads)
    return vals, vld
