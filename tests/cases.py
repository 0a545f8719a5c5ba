"""Expression catalogue + data generators shared by the CPU tests (oracle, NVRTC compile) and
the GPU parity tests.  Each case is a function of a TreeExprBuilder-like `b` returning
(schema, [(root_node, result_type)], kind) where kind is "project" or "filter"."""
from __future__ import annotations

import decimal

import numpy as np
import pyarrow as pa

INT_TYPES = [pa.int8(), pa.int16(), pa.int32(), pa.int64(), pa.uint8(), pa.uint16(), pa.uint32(),
             pa.uint64()]
FLOAT_TYPES = [pa.float32(), pa.float64()]
NUMERIC = INT_TYPES + FLOAT_TYPES
RELOPS = ["equal", "not_equal", "less_than", "less_than_or_equal_to", "greater_than",
          "greater_than_or_equal_to"]

_NP = {pa.int8(): np.int8, pa.int16(): np.int16, pa.int32(): np.int32, pa.int64(): np.int64,
       pa.uint8(): np.uint8, pa.uint16(): np.uint16, pa.uint32(): np.uint32,
       pa.uint64(): np.uint64, pa.float32(): np.float32, pa.float64(): np.float64}

WORDS = ["special", "requests", "spark", "park", "fire", "carefully", "final", "deposits", "ironic",
         "blithely", "Quick", "BROWN", "fox", "a", "", "x_y", "100%", "naïve", "日本語", "ünï", "  pad  ",
         "SPECIAL REQUESTS", "special packages requests"]


def random_array(t: pa.DataType, n: int, rng: np.random.Generator, null_prob: float = 0.1,
                 small: bool = False) -> pa.Array:
    """Random Arrow array of type t with edge values sprinkled in and ~null_prob nulls."""
    mask = rng.random(n) < null_prob if null_prob > 0 else None
    if t in _NP:
        npt = _NP[t]
        if t in FLOAT_TYPES:
            vals = (rng.standard_normal(n) * (10.0 if small else 1e6)).astype(npt)
            edge = np.array([0.0, -0.0, 1.0, -1.0, 0.5, 1e-30, 3.5, 1000.0], dtype=npt)
        else:
            info = np.iinfo(npt)
            if small:
                lo, hi = max(info.min, -50), min(info.max, 50)
                vals = rng.integers(lo, hi + 1, n, dtype=np.int64).astype(npt)
                edge = np.array([0, 1, 2, 3], dtype=npt)
            else:
                vals = rng.integers(info.min, int(info.max) + 1, n, dtype=npt)
                edge = np.array([info.min, info.max, 0, 1, info.max - 1, info.min + 1], dtype=npt)
        k = min(n, len(edge))
        if k:
            pos = rng.choice(n, k, replace=False)
            vals[pos] = edge[:k]
        return pa.array(vals, type=t, mask=mask)
    if pa.types.is_boolean(t):
        return pa.array(rng.random(n) < 0.5, type=t, mask=mask)
    if pa.types.is_date32(t):
        return pa.array(rng.integers(-30000, 60000, n).astype(np.int32), type=pa.int32(), mask=mask).cast(t)
    if pa.types.is_date64(t):
        days = rng.integers(-30000, 60000, n).astype(np.int64)
        return pa.array(days * 86400000, type=pa.int64(), mask=mask).cast(t)
    if pa.types.is_timestamp(t):
        ms = rng.integers(-2_000_000_000_000, 4_000_000_000_000, n).astype(np.int64)
        return pa.array(ms, type=pa.int64(), mask=mask).cast(t)
    if pa.types.is_time32(t):
        return pa.array(rng.integers(0, 86400000, n).astype(np.int32), type=pa.int32(), mask=mask).cast(t)
    if pa.types.is_decimal128(t):
        digits = t.precision if not small else min(t.precision, 6)
        out = []
        for i in range(n):
            if mask is not None and mask[i]:
                out.append(None)
                continue
            nd = int(rng.integers(1, digits + 1))
            mag = int("".join(str(int(d)) for d in rng.integers(0, 10, nd)))
            if rng.random() < 0.5:
                mag = -mag
            out.append(decimal.Decimal(mag).scaleb(-t.scale))
        return pa.array(out, type=t)
    if pa.types.is_string(t) or pa.types.is_binary(t):
        out = []
        for i in range(n):
            if mask is not None and mask[i]:
                out.append(None)
                continue
            k = int(rng.integers(0, 5))
            s = " ".join(WORDS[int(j)] for j in rng.integers(0, len(WORDS), k))
            out.append(s if pa.types.is_string(t) else s.encode("utf-8"))
        return pa.array(out, type=t)
    raise NotImplementedError(str(t))


def random_batch(schema: pa.Schema, n: int, seed: int, null_prob: float = 0.1, offset: int = 0,
                 small: bool = False) -> pa.RecordBatch:
    """Random batch; with offset > 0 every column is a slice (ArrayData.offset != 0)."""
    rng = np.random.default_rng(seed)
    cols = []
    for f in schema:
        arr = random_array(f.type, n + offset, rng, null_prob, small)
        cols.append(arr.slice(offset) if offset else arr)
    return pa.RecordBatch.from_arrays(cols, schema=schema)


# ---- expression cases ----------------------------------------------------------------------
def F(b, name, t):
    return b.make_field(pa.field(name, t))


def case_arith(op, t):
    def build(b):
        schema = pa.schema([("a", t), ("b", t)])
        return schema, [(b.make_function(op, [F(b, "a", t), F(b, "b", t)], t), t)], "project"
    build.__name__ = "%s_%s" % (op, t)
    return build


def case_relop(op, t):
    def build(b):
        schema = pa.schema([("a", t), ("b", t)])
        return schema, [(b.make_function(op, [F(b, "a", t), F(b, "b", t)], pa.bool_()), pa.bool_())], "project"
    build.__name__ = "%s_%s" % (op, t)
    return build


def case_if_else(b):
    t = pa.int32()
    schema = pa.schema([("a", t), ("b", t), ("c", t)])
    a, bb, c = F(b, "a", t), F(b, "b", t), F(b, "c", t)
    cond1 = b.make_function("greater_than", [a, bb], pa.bool_())
    cond2 = b.make_function("less_than", [bb, c], pa.bool_())
    inner = b.make_if(cond2, bb, b.make_function("add", [c, b.make_literal(7, t)], t), t)
    root = b.make_if(cond1, a, inner, t)
    return schema, [(root, t)], "project"


def case_if_null_literal(b):
    t = pa.int64()
    schema = pa.schema([("a", t), ("b", t)])
    a, bb = F(b, "a", t), F(b, "b", t)
    cond = b.make_function("less_than", [a, bb], pa.bool_())
    root = b.make_if(cond, a, b.make_literal(None, t), t)
    return schema, [(root, t)], "project"


def case_kleene(b):
    t = pa.bool_()
    schema = pa.schema([("x", t), ("y", t), ("z", t)])
    x, y, z = F(b, "x", t), F(b, "y", t), F(b, "z", t)
    return schema, [(b.make_and([x, y]), t), (b.make_or([x, y]), t), (b.make_and([x, y, z]), t),
                    (b.make_or([b.make_and([x, y]), b.make_function("not", [z], t)]), t)], "project"


def case_null_tests(b):
    t = pa.float64()
    schema = pa.schema([("a", t), ("b", t), ("x", pa.bool_())])
    a, bb, x = F(b, "a", t), F(b, "b", t), F(b, "x", pa.bool_())
    B = pa.bool_()
    return schema, [(b.make_function("isnull", [a], B), B), (b.make_function("isnotnull", [bb], B), B),
                    (b.make_function("is_distinct_from", [a, bb], B), B),
                    (b.make_function("is_not_distinct_from", [a, bb], B), B),
                    (b.make_function("istrue", [x], B), B), (b.make_function("isnotfalse", [x], B), B)], "project"


def case_casts(b):
    schema = pa.schema([("i", pa.int32()), ("l", pa.int64()), ("f", pa.float32()), ("d", pa.float64())])
    i, l, f, d = F(b, "i", pa.int32()), F(b, "l", pa.int64()), F(b, "f", pa.float32()), F(b, "d", pa.float64())
    return schema, [(b.make_function("castBIGINT", [i], pa.int64()), pa.int64()),
                    (b.make_function("castINT", [l], pa.int32()), pa.int32()),
                    (b.make_function("castFLOAT4", [l], pa.float32()), pa.float32()),
                    (b.make_function("castFLOAT4", [d], pa.float32()), pa.float32()),
                    (b.make_function("castFLOAT8", [l], pa.float64()), pa.float64()),
                    (b.make_function("castFLOAT8", [f], pa.float64()), pa.float64()),
                    (b.make_function("abs", [i], pa.int32()), pa.int32()),
                    (b.make_function("negative", [d], pa.float64()), pa.float64()),
                    (b.make_function("bitwise_xor", [l, l], pa.int64()), pa.int64())], "project"


def case_mod(b):
    schema = pa.schema([("l", pa.int64()), ("i", pa.int32()), ("m", pa.int64())])
    l, i, m = F(b, "l", pa.int64()), F(b, "i", pa.int32()), F(b, "m", pa.int64())
    return schema, [(b.make_function("mod", [l, i], pa.int32()), pa.int32()),
                    (b.make_function("mod", [l, m], pa.int64()), pa.int64())], "project"


def case_dates(b):
    ts, d64, d32 = pa.timestamp("ms"), pa.date64(), pa.date32()
    schema = pa.schema([("t", ts), ("d", d64), ("e", d32)])
    t, d, e = F(b, "t", ts), F(b, "d", d64), F(b, "e", d32)
    L = pa.int64()
    outs = [(b.make_function(fn, [t], L), L) for fn in
            ["extractYear", "extractMonth", "extractDay", "extractHour", "extractMinute",
             "extractSecond", "extractDoy", "extractDow", "extractQuarter", "extractEpoch"]]
    outs += [(b.make_function("extractYear", [d], L), L), (b.make_function("extractMonth", [e], L), L),
             (b.make_function("extractDay", [e], L), L),
             (b.make_function("castDATE", [t], d64), d64),
             (b.make_function("less_than", [e, b.make_literal(9131, d32)], pa.bool_()), pa.bool_())]
    return schema, outs, "project"


def case_rounding(b):
    schema = pa.schema([("d", pa.float64()), ("f", pa.float32()), ("s", pa.int32()), ("l", pa.int64())])
    d, f, s, l = F(b, "d", pa.float64()), F(b, "f", pa.float32()), F(b, "s", pa.int32()), F(b, "l", pa.int64())
    D = pa.float64()
    outs = [(b.make_function("round", [d], D), D), (b.make_function("round", [f], pa.float32()), pa.float32()),
            (b.make_function("ceil", [d], D), D), (b.make_function("floor", [d], D), D),
            (b.make_function("truncate", [d], D), D), (b.make_function("round", [l], pa.int64()), pa.int64())]
    for k in (0, 1, 2, 5, -1, -3):
        outs.append((b.make_function("round", [d, b.make_literal(k, pa.int32())], D), D))
    outs.append((b.make_function("round", [d, s], D), D))
    return schema, outs, "project"


def case_intmath(b):
    """div / pmod / sign / greatest / least / nvl / round and truncate with a scale."""
    I, L, F32, D, B = pa.int32(), pa.int64(), pa.float32(), pa.float64(), pa.bool_()
    schema = pa.schema([("i", I), ("j", I), ("l", L), ("m", L), ("f", F32), ("g", F32), ("d", D), ("e", D)])
    i, j, l, m = F(b, "i", I), F(b, "j", I), F(b, "l", L), F(b, "m", L)
    f, g, d, e = F(b, "f", F32), F(b, "g", F32), F(b, "d", D), F(b, "e", D)
    fn = b.make_function
    lit = b.make_literal
    outs = [
        (b.make_if(fn("not_equal", [j, lit(0, I)], B), fn("div", [i, j], I), i, I), I),
        (b.make_if(fn("not_equal", [m, lit(0, L)], B), fn("div", [l, m], L), l, L), L),
        (fn("pmod", [i, j], I), I), (fn("pmod", [l, m], L), L),
        (fn("pmod", [i, lit(7, I)], I), I), (fn("pmod", [l, lit(-7, L)], L), L),
        (fn("sign", [i], I), I), (fn("sign", [l], L), L), (fn("sign", [f], F32), F32), (fn("sign", [d], D), D),
        (fn("greatest", [i, j], I), I), (fn("least", [i, j, lit(0, I)], I), I),
        (fn("greatest", [l, m, lit(5, L), l], L), L), (fn("least", [l, m], L), L),
        (fn("greatest", [f, g], F32), F32), (fn("least", [d, e, lit(0.5, D)], D), D),
        (fn("greatest", [d, e], D), D), (fn("least", [f, g, f, g], F32), F32),
        (fn("nvl", [i, j], I), I), (fn("nvl", [d, lit(-1.5, D)], D), D), (fn("nvl", [l, fn("add", [m, m], L)], L), L),
        (fn("round", [i, lit(-2, I)], I), I), (fn("round", [l, lit(-5, I)], L), L), (fn("round", [l, j], L), L),
        (fn("round", [i, lit(3, I)], I), I), (fn("round", [l, lit(-19, I)], L), L), (fn("round", [l, lit(-40, I)], L), L),
        (fn("truncate", [l, lit(-3, I)], L), L), (fn("truncate", [i, lit(-1, I)], I), I),
        (fn("truncate", [d, lit(2, I)], D), D), (fn("truncate", [d, lit(-2, I)], D), D), (fn("trunc", [d, j], D), D),
        (fn("bround", [d], D), D), (fn("bround", [fn("divide", [fn("castFLOAT8", [i], D), lit(2.0, D)], D)], D), D),
        # factorial raises outside 0..20: pmod(i, 21) keeps it inside
        (fn("factorial", [fn("pmod", [i, lit(21, I)], I)], L), L), (fn("factorial", [fn("pmod", [l, lit(21, L)], L)], L), L),
    ]
    return schema, outs, "project"


def case_calendar(b):
    """ISO week, decade / century / millennium, date_trunc_*, last_day, time of day."""
    ts, d64, t32, L = pa.timestamp("ms"), pa.date64(), pa.time32("ms"), pa.int64()
    schema = pa.schema([("t", ts), ("d", d64), ("c", t32)])
    t, d, c = F(b, "t", ts), F(b, "d", d64), F(b, "c", t32)
    fn = b.make_function
    outs = []
    for name in ("extractWeek", "extractDecade", "extractCentury", "extractMillennium"):
        outs.append((fn(name, [t], L), L))
        outs.append((fn(name, [d], L), L))
    for unit in ("Second", "Minute", "Hour", "Day", "Week", "Month", "Quarter", "Year", "Decade", "Century",
                 "Millennium"):
        outs.append((fn("date_trunc_" + unit, [t], ts), ts))
    outs.append((fn("date_trunc_Month", [d], d64), d64))
    outs.append((fn("date_trunc_Week", [d], d64), d64))
    outs.append((fn("last_day", [t], d64), d64))
    outs.append((fn("last_day", [d], d64), d64))
    outs.append((fn("castTIME", [t], t32), t32))
    for name in ("extractHour", "extractMinute", "extractSecond"):
        outs.append((fn(name, [c], L), L))
    outs.append((fn("extractHour", [fn("castTIME", [t], t32)], L), L))
    return schema, outs, "project"


def case_string_positions(b):
    """ascii / left / right / locate / strpos / byte_substr / ilike / nvl over strings."""
    S, BIN, I, B = pa.string(), pa.binary(), pa.int32(), pa.bool_()
    schema = pa.schema([("s", S), ("u", S), ("z", BIN), ("k", I)])
    s, u, z, k = F(b, "s", S), F(b, "u", S), F(b, "z", BIN), F(b, "k", I)
    fn = b.make_function
    lit = b.make_literal
    clen = lambda x: fn("char_length", [x], I)
    kk = fn("subtract", [fn("castINT", [fn("mod", [fn("castBIGINT", [k], pa.int64()), lit(9, pa.int64())], pa.int64())], I),
                         lit(4, I)], I)   # k mod 9 - 4: a small signed count
    outs = [
        (fn("ascii", [s], I), I), (fn("ascii", [fn("upper", [s], S)], I), I),
        (clen(fn("left", [s, lit(3, I)], S)), I), (clen(fn("left", [s, lit(-2, I)], S)), I),
        (clen(fn("right", [s, lit(4, I)], S)), I), (clen(fn("right", [s, lit(-3, I)], S)), I),
        (fn("equal", [fn("left", [s, kk], S), fn("right", [u, kk], S)], B), B),
        (fn("octet_length", [fn("right", [s, kk], S)], I), I),
        (fn("locate", [lit("ar", S), s], I), I), (fn("locate", [lit("", S), s], I), I),
        (fn("locate", [lit("本", S), s], I), I), (fn("locate", [lit("e", S), s, lit(3, I)], I), I),
        (fn("locate", [lit("s", S), s, b.make_if(fn("greater_than", [kk, lit(0, I)], B), kk, lit(1, I), I)], I), I),
        (fn("locate", [u, s], I), I),
        (fn("position", [lit("re", S), s], I), I),
        (fn("strpos", [s, lit("re", S)], I), I), (fn("strpos", [fn("upper", [s], S), lit("RE", S)], I), I),
        (fn("octet_length", [fn("byte_substr", [z, lit(2, I), lit(5, I)], BIN)], I), I),
        (fn("octet_length", [fn("byte_substr", [z, lit(-3, I), lit(2, I)], BIN)], I), I),
        (fn("equal", [fn("byte_substr", [z, kk, lit(3, I)], BIN), fn("byte_substr", [z, lit(1, I), lit(3, I)], BIN)], B), B),
        (fn("ilike", [s, lit("%SPecial%Requests%", S)], B), B), (fn("ilike", [s, lit("quick%", S)], B), B),
        (fn("ilike", [fn("substr", [s, lit(1, pa.int64()), lit(12, pa.int64())], S), lit("%BROWN%", S)], B), B),
        (clen(fn("nvl", [s, u], S)), I), (fn("equal", [fn("nvl", [s, lit("none", S)], S), u], B), B),
    ]
    return schema, outs, "project"


def case_number_to_text(b):
    """castVARCHAR of integers / booleans / dates / timestamps: text written into a thread-private
    scratch slot, consumed per lane (lengths, LIKE, compares, hashes) or projected (alone, inside a
    concat rope, chosen by if/else)."""
    S, I, L, B, TS, D64 = pa.string(), pa.int32(), pa.int64(), pa.bool_(), pa.timestamp("ms"), pa.date64()
    schema = pa.schema([("i", I), ("l", L), ("p", B), ("t", TS), ("w", D64), ("s", S)])
    i, l, pp, t, w, s = F(b, "i", I), F(b, "l", L), F(b, "p", B), F(b, "t", TS), F(b, "w", D64), F(b, "s", S)
    fn = b.make_function
    n = lambda v: b.make_literal(v, L)
    txt = lambda x, k: fn("castVARCHAR", [x, n(k)], S)
    outs = [
        (txt(i, 20), S), (txt(l, 30), S), (txt(l, 5), S), (txt(l, 0), S), (txt(pp, 10), S), (txt(pp, 3), S),
        (txt(t, 30), S), (txt(t, 16), S), (txt(w, 10), S), (txt(w, 40), S),
        (fn("concat", [b.make_literal("id-", S), txt(l, 30), b.make_literal("/", S), txt(i, 4)], S), S),
        (fn("concatOperator", [s, txt(pp, 10)], S), S),
        (b.make_if(pp, txt(i, 12), fn("upper", [s], S), S), S),
        (fn("upper", [txt(pp, 10)], S), S),
        (fn("char_length", [txt(l, 30)], I), I), (fn("octet_length", [txt(t, 30)], I), I),
        (fn("like", [txt(l, 30), b.make_literal("%12%", S)], B), B),
        (fn("equal", [txt(i, 12), fn("castVARCHAR", [s, n(12)], S)], B), B),
        (fn("hash32", [txt(l, 30)], I), I),
        (fn("castBIGINT", [txt(l, 30)], L), L),   # text -> number again: the round trip is the identity
        (fn("castFLOAT8", [txt(l, 30)], pa.float64()), pa.float64()),   # == castFLOAT8(l) up to 2^53, RNE beyond
        (fn("castFLOAT4", [txt(i, 12)], pa.float32()), pa.float32()),
        (fn("substr", [txt(t, 30), n(12), n(8)], S), S),
        (txt(fn("castFLOAT8", [l], pa.float64()), 30), S), (txt(fn("castFLOAT4", [i], pa.float32()), 30), S),
        (txt(fn("divide", [fn("castFLOAT8", [l], pa.float64()), b.make_literal(3.0e9, pa.float64())], pa.float64()), 12), S),
    ]
    return schema, outs, "project"


def case_string_misc(b):
    """trim with a character set, split_part, crc32, to_hex, degrees / radians, datediff."""
    S, BIN, I, L, B, D, TS, D64 = (pa.string(), pa.binary(), pa.int32(), pa.int64(), pa.bool_(), pa.float64(),
                                   pa.timestamp("ms"), pa.date64())
    schema = pa.schema([("s", S), ("u", S), ("z", BIN), ("i", I), ("l", L), ("d", D), ("t", TS), ("v", TS), ("w", D64)])
    s, u, z, i, l, d, t, v, w = (F(b, n, ty) for n, ty in zip("suzildtvw", (S, S, BIN, I, L, D, TS, TS, D64)))
    fn = b.make_function
    lit = b.make_literal
    olen = lambda x: fn("octet_length", [x], I)
    kk = fn("add", [fn("castINT", [fn("pmod", [l, lit(4, L)], L)], I), lit(1, I)], I)   # 1..4
    outs = [
        (fn("ltrim", [s, lit("sp ", S)], S), S), (fn("rtrim", [s, lit("se ", S)], S), S),
        (fn("btrim", [s, lit(" ü日", S)], S), S), (olen(fn("trim", [s, u], S)), I),
        (fn("split_part", [s, lit(" ", S), lit(1, I)], S), S), (fn("split_part", [s, lit(" ", S), lit(3, I)], S), S),
        (olen(fn("split_part", [s, lit("e", S), kk], S)), I), (fn("split_part", [s, lit("", S), lit(1, I)], S), S),
        (fn("upper", [fn("split_part", [s, lit("re", S), lit(2, I)], S)], S), S),
        (fn("crc32", [s], L), L), (fn("crc32", [z], L), L), (fn("crc32", [fn("upper", [s], S)], L), L),
        (fn("to_hex", [l], S), S), (fn("to_hex", [i], S), S),
        (fn("concat", [lit("0x", S), fn("to_hex", [i], S)], S), S),
        (fn("degrees", [d], D), D), (fn("radians", [d], D), D),
        (fn("datediff", [t, v], I), I), (fn("datediff", [w, fn("castDATE", [t], D64)], I), I),
    ]
    return schema, outs, "project"


def case_digests(b):
    """hashSHA256 / hashSHA1 / hashMD5 of utf8 and binary (lower-case hex), alone and consumed."""
    S, BIN, I, B = pa.string(), pa.binary(), pa.int32(), pa.bool_()
    schema = pa.schema([("s", S), ("z", BIN)])
    s, z = F(b, "s", S), F(b, "z", BIN)
    fn = b.make_function
    outs = [(fn("hashSHA256", [s], S), S), (fn("hashSHA1", [s], S), S), (fn("hashMD5", [s], S), S),
            (fn("sha256", [z], S), S), (fn("sha1", [z], S), S), (fn("md5", [z], S), S),
            (fn("hashSHA256", [fn("upper", [s], S)], S), S),
            (fn("concat", [fn("md5", [s], S), b.make_literal(":", S), fn("sha1", [s], S)], S), S),
            (fn("like", [fn("hashMD5", [s], S), b.make_literal("%a%f%", S)], B), B),
            (fn("char_length", [fn("hashSHA256", [z], S)], I), I)]
    return schema, outs, "project"


def digest_batch(n: int, seed: int, offset: int = 0) -> pa.RecordBatch:
    """Messages around the 55 / 56 / 64-byte padding boundaries, multi-block ones, empty, nulls."""
    rng = np.random.default_rng(seed)
    lens = [0, 1, 3, 55, 56, 57, 63, 64, 65, 111, 112, 119, 120, 128, 200]
    ss, zz = [], []
    for k in range(n + offset):
        ln = lens[k % len(lens)] if k < 4 * len(lens) else int(rng.integers(0, 180))
        raw = bytes(rng.integers(32, 127, ln).astype(np.uint8))
        ss.append(None if k % 17 == 11 else raw.decode("ascii"))
        zz.append(None if k % 19 == 7 else bytes(rng.integers(0, 256, ln).astype(np.uint8)))
    ss[0] = "abc"
    a, zc = pa.array(ss, pa.string()), pa.array(zz, pa.binary())
    if offset:
        a, zc = a.slice(offset), zc.slice(offset)
    return pa.RecordBatch.from_arrays([a, zc], names=["s", "z"])


def case_virtual_strings(b):
    """repeat / space / reverse / lpad / rpad: periodic and reversed pieces that only the string write
    pass reads; alone, inside concat ropes, chosen by if/else, over scratch-slot text."""
    S, I, L, B = pa.string(), pa.int32(), pa.int64(), pa.bool_()
    schema = pa.schema([("s", S), ("u", S), ("k", I), ("l", L), ("p", B)])
    s, u, k, l, pp = F(b, "s", S), F(b, "u", S), F(b, "k", I), F(b, "l", L), F(b, "p", B)
    fn = b.make_function
    lit = b.make_literal
    kk = fn("subtract", [fn("castINT", [fn("pmod", [fn("castBIGINT", [k], L), lit(23, L)], L)], I), lit(3, I)], I)   # -3..19
    outs = [
        (fn("repeat", [s, lit(3, I)], S), S), (fn("repeat", [fn("upper", [s], S), kk], S), S),
        (fn("space", [kk], S), S), (fn("reverse", [s], S), S), (fn("reverse", [fn("lower", [u], S)], S), S),
        (fn("lpad", [s, lit(12, I)], S), S), (fn("rpad", [s, lit(12, I), lit("*", S)], S), S),
        (fn("lpad", [s, kk, lit("ab", S)], S), S), (fn("rpad", [s, kk, lit("日本x", S)], S), S),
        (fn("lpad", [s, lit(7, I), u], S), S), (fn("rpad", [u, lit(9, I), lit("", S)], S), S),
        (fn("concat", [lit("[", S), fn("lpad", [fn("castVARCHAR", [l, lit(30, L)], S), lit(8, I), lit("0", S)], S), lit("]", S)], S), S),
        (fn("concat", [fn("reverse", [s], S), fn("space", [lit(2, I)], S), fn("repeat", [lit("-", S), lit(4, I)], S)], S), S),
        (b.make_if(pp, fn("repeat", [u, lit(2, I)], S), fn("reverse", [u], S), S), S),
        (fn("reverse", [fn("castVARCHAR", [l, lit(30, L)], S)], S), S),
        (fn("repeat", [fn("castVARCHAR", [k, lit(4, L)], S), lit(2, I)], S), S),
        (fn("replace", [s, lit("re", S), lit("<RE>", S)], S), S), (fn("replace", [s, lit(" ", S), lit("", S)], S), S),
        (fn("replace", [fn("upper", [s], S), lit("SPECIAL", S), lit("日本", S)], S), S),
        (fn("replace", [u, lit("", S), lit("x", S)], S), S), (fn("replace", [u, lit("ss", S), lit("s", S)], S), S),
        (fn("concat", [fn("replace", [s, lit("a", S), lit("aa", S)], S), lit("|", S),
                       fn("replace", [fn("castVARCHAR", [l, lit(30, L)], S), lit("1", S), lit("one", S)], S)], S), S),
        (b.make_if(pp, fn("replace", [s, lit("e", S), lit("", S)], S), fn("replace", [s, lit("e", S), lit("EE", S)], S), S), S),
    ]
    return schema, outs, "project"


def case_math(b):
    """exp / log / ln / log10 / cbrt (explicit IEEE sequences, bit-exact against the oracle)."""
    D = pa.float64()
    schema = pa.schema([("d", D), ("e", D)])
    d, e = F(b, "d", D), F(b, "e", D)
    fn = b.make_function
    small = fn("divide", [d, b.make_literal(1.0e5, D)], D)            # ~N(0, 10): exp stays finite
    outs = [(fn("exp", [small], D), D), (fn("exp", [d], D), D), (fn("log", [fn("abs", [d], D)], D), D),
            (fn("ln", [e], D), D), (fn("log10", [fn("abs", [e], D)], D), D), (fn("cbrt", [d], D), D),
            (fn("log", [fn("exp", [small], D)], D), D), (fn("cbrt", [fn("multiply", [e, fn("multiply", [e, e], D)], D)], D), D)]
    return schema, outs, "project"


def case_trig(b):
    """sin / cos / tan / cot of doubles (ordinary, tiny, huge: the integer Payne-Hanek reduction), floats and
    integers; bit-exact against the oracle."""
    D, F4, I, L = pa.float64(), pa.float32(), pa.int32(), pa.int64()
    schema = pa.schema([("d", D), ("f", F4), ("i", I), ("l", L)])
    d, f, i, l = F(b, "d", D), F(b, "f", F4), F(b, "i", I), F(b, "l", L)
    fn = b.make_function
    small = fn("divide", [d, b.make_literal(1.0e6, D)], D)
    tiny = fn("divide", [d, b.make_literal(1.0e20, D)], D)
    huge = fn("multiply", [fn("multiply", [d, b.make_literal(1.0e150, D)], D), b.make_literal(1.0e140, D)], D)
    outs = []
    for name in ("sin", "cos", "tan", "cot"):
        outs += [(fn(name, [x], D), D) for x in (d, small, tiny, huge, f, i, l)]
    outs.append((fn("add", [fn("multiply", [fn("sin", [d], D), fn("sin", [d], D)], D),
                            fn("multiply", [fn("cos", [d], D), fn("cos", [d], D)], D)], D), D))
    return schema, outs, "project"


def case_regexp(b):
    """regexp_matches / regexp_like on the generic string columns (tests/test_parity_gpu.py::test_regexp_matches
    holds the pattern matrix)."""
    S, B = pa.string(), pa.bool_()
    schema = pa.schema([("s", S), ("u", S)])
    s, u = F(b, "s", S), F(b, "u", S)
    fn = b.make_function
    lit = lambda v: b.make_literal(v, S)
    outs = [(fn("regexp_matches", [s, lit(p)], B), B) for p in
            ("^[a-z]+$", "[0-9]{2,}", "(spec|requ).*s$", "^\\s*$", "[^ -~]", "a.e", "\\w+\\s\\w+")]
    outs.append((fn("regexp_like", [fn("upper", [u], S), lit("^[A-Z ]*$")], B), B))
    outs.append((b.make_if(fn("regexp_matches", [s, lit("e")], B), fn("char_length", [s], pa.int32()),
                           b.make_literal(-1, pa.int32()), pa.int32()), pa.int32()))
    return schema, outs, "project"


def case_misc_casts(b):
    """castINT / castBIGINT of floats (round half away, saturating), to_timestamp / to_time, find_in_set, instr."""
    D, F4, I, L, S, TS, T32 = pa.float64(), pa.float32(), pa.int32(), pa.int64(), pa.string(), pa.timestamp("ms"), pa.time32("ms")
    schema = pa.schema([("d", D), ("f", F4), ("i", I), ("l", L), ("s", S), ("u", S)])
    d, f, i, l, s, u = (F(b, n, t) for n, t in zip("dfilsu", (D, F4, I, L, S, S)))
    fn = b.make_function
    half = fn("add", [fn("divide", [d, b.make_literal(2.0, D)], D), b.make_literal(0.5, D)], D)
    huge = fn("multiply", [d, b.make_literal(1.0e13, D)], D)
    outs = [(fn("castBIGINT", [d], L), L), (fn("castBIGINT", [half], L), L), (fn("castBIGINT", [huge], L), L), (fn("castBIGINT", [f], L), L),
            (fn("castINT", [d], I), I), (fn("castINT", [half], I), I), (fn("castINT", [huge], I), I), (fn("castINT", [f], I), I),
            (fn("castINT", [fn("sqrt", [d], D)], I), I),   # NaN for negative d -> 0
            (fn("to_timestamp", [i], TS), TS), (fn("to_timestamp", [l], TS), TS), (fn("to_timestamp", [d], TS), TS),
            (fn("to_timestamp", [f], TS), TS), (fn("to_time", [i], T32), T32), (fn("to_time", [l], T32), T32),
            (fn("to_time", [d], T32), T32), (fn("to_time", [f], T32), T32),
            (fn("find_in_set", [s, b.make_literal("fox,special,,requests,日本語,the", S)], I), I),
            (fn("find_in_set", [b.make_literal("", S), u], I), I), (fn("find_in_set", [s, u], I), I),
            (fn("find_in_set", [fn("lower", [s], S), fn("lower", [u], S)], I), I),
            (fn("instr", [s, b.make_literal("e", S)], I), I), (fn("instr", [u, s], I), I)]
    return schema, outs, "project"


def case_power(b):
    """power / pow: ordinary, huge, tiny and negative bases, integer and fractional exponents."""
    D = pa.float64()
    schema = pa.schema([("d", D), ("e", D)])
    d, e = F(b, "d", D), F(b, "e", D)
    fn = b.make_function
    lit = lambda v: b.make_literal(v, D)
    small = fn("divide", [d, lit(2.0e5)], D)
    near1 = fn("add", [lit(1.0), fn("divide", [e, lit(1.0e18)], D)], D)
    outs = [(fn("power", [fn("abs", [d], D), fn("divide", [e, lit(1.0e6)], D)], D), D),
            (fn("power", [d, fn("round", [fn("divide", [e, lit(3.0e5)], D)], D)], D), D),       # negative bases, integer exponents
            (fn("power", [d, small], D), D),                                                         # mostly NaN / huge
            (fn("pow", [near1, lit(1.0e15)], D), D), (fn("pow", [lit(2.0), small], D), D),
            (fn("power", [fn("abs", [small], D), lit(0.5)], D), D), (fn("power", [d, lit(2.0)], D), D),
            (fn("power", [d, lit(-3.0)], D), D), (fn("power", [lit(10.0), fn("divide", [e, lit(4.0e3)], D)], D), D)]
    nz = fn("add", [fn("abs", [e], D), lit(0.25)], D)
    outs += [(fn("mod", [d, nz], D), D), (fn("modulo", [fn("multiply", [d, lit(1.0e200)], D), lit(3.0)], D), D),
             (fn("mod", [lit(5.5), fn("subtract", [lit(0.0), nz], D)], D), D)]
    outs += [(fn("log", [lit(10.0), fn("abs", [d], D)], D), D), (fn("log", [fn("add", [fn("abs", [e], D), lit(2.0)], D), fn("abs", [d], D)], D), D)]
    tiny = fn("divide", [d, lit(1.0e14)], D)
    for name in ("sinh", "cosh", "tanh"):
        outs += [(fn(name, [x], D), D) for x in (small, tiny, fn("divide", [d, lit(1.5e3)], D), d)]
    return schema, outs, "project"


def case_inverse_trig(b):
    """atan / atan2 / asin / acos: ordinary, tiny and huge ratios, all quadrants, |v| <= 1 and beyond (NaN)."""
    D = pa.float64()
    schema = pa.schema([("d", D), ("e", D)])
    d, e = F(b, "d", D), F(b, "e", D)
    fn = b.make_function
    lit = lambda v: b.make_literal(v, D)
    unit = fn("sin", [d], D)                                  # in [-1, 1]
    outs = [(fn("atan2", [d, e], D), D), (fn("atan2", [d, fn("multiply", [e, lit(1.0e12)], D)], D), D),
            (fn("atan2", [fn("multiply", [d, lit(1.0e25)], D), e], D), D), (fn("atan2", [d, lit(0.0)], D), D),
            (fn("atan2", [lit(0.0), e], D), D), (fn("atan", [d], D), D), (fn("atan", [fn("divide", [d, lit(1.0e17)], D)], D), D),
            (fn("asin", [unit], D), D), (fn("acos", [unit], D), D), (fn("asin", [fn("divide", [d, lit(1.0e6)], D)], D), D),
            (fn("acos", [fn("divide", [d, lit(1.0e6)], D)], D), D), (fn("acos", [fn("divide", [unit, lit(1.0e12)], D)], D), D),
            (fn("asin", [fn("cos", [fn("divide", [d, lit(1.0e14)], D)], D)], D), D)]       # next to 1
    return schema, outs, "project"


def case_in_floats(b):
    """IN over float64 / float32 (C++ MakeInExpressionDouble / Float): IEEE equality, -0.0 is in {0.0}, NaN in nothing."""
    D, F4, B = pa.float64(), pa.float32(), pa.bool_()
    schema = pa.schema([("d", D), ("f", F4)])
    d, f = F(b, "d", D), F(b, "f", F4)
    fn = b.make_function
    small = fn("round", [fn("divide", [d, b.make_literal(2.0e5, D)], D)], D)      # small integers, some -0.0
    outs = [(b.make_in_expression(small, [0.0, 1.0, -2.0, 3.5, float("nan")], D), B),
            (b.make_in_expression(d, [], D), B),
            (b.make_in_expression(fn("castFLOAT4", [small], F4), [-0.0, 2.0, 5.0, 1e30, -1.0, 7.0, 8.0, 9.0, 10.0, 11.0], F4), B),
            (b.make_in_expression(fn("multiply", [small, b.make_literal(0.5, D)], D), [0.5, -0.5, 1.5], D), B)]
    return schema, outs, "project"


def case_concat_consumers(b):
    """upper / lower and the length functions over concat / concatOperator: rewritten to act on the pieces
    (the oracle evaluates the tree as written)."""
    S, I = pa.string(), pa.int32()
    schema = pa.schema([("s", S), ("u", S)])
    s, u = F(b, "s", S), F(b, "u", S)
    fn = b.make_function
    lit = lambda v: b.make_literal(v, S)
    cc = fn("concat", [s, lit(" - "), u], S)
    co = fn("concatOperator", [s, u], S)
    nested = fn("concat", [fn("upper", [co], S), lit("日本"), fn("lower", [cc], S)], S)
    outs = [(fn("upper", [cc], S), S), (fn("lower", [co], S), S), (fn("char_length", [cc], I), I), (fn("octet_length", [co], I), I),
            (fn("bit_length", [nested], I), I), (fn("length", [fn("upper", [cc], S)], I), I), (nested, S),
            (fn("char_length", [fn("concat", [lit("é"), s], S)], I), I),
            (b.make_if(fn("greater_than", [fn("char_length", [co], I), b.make_literal(60, I)], pa.bool_()), fn("upper", [co], S), cc, S), S)]
    return schema, outs, "project"


def case_date_arith(b):
    ts, d64 = pa.timestamp("ms"), pa.date64()
    schema = pa.schema([("t", ts), ("u", ts), ("d", d64), ("n", pa.int32()), ("m", pa.int64())])
    t, u, d, n, m = F(b, "t", ts), F(b, "u", ts), F(b, "d", d64), F(b, "n", pa.int32()), F(b, "m", pa.int64())
    outs = []
    for fn in ("timestampaddSecond", "timestampaddMinute", "timestampaddHour", "timestampaddDay", "timestampaddWeek",
               "timestampaddMonth", "timestampaddQuarter", "timestampaddYear"):
        outs.append((b.make_function(fn, [n, t], ts), ts))
    outs.append((b.make_function("timestampaddDay", [m, t], ts), ts))
    outs.append((b.make_function("timestampaddMonth", [b.make_literal(1, pa.int32()), t], ts), ts))
    outs.append((b.make_function("date_add", [d, n], d64), d64))
    outs.append((b.make_function("date_sub", [d, n], d64), d64))
    outs.append((b.make_function("date_add", [t, n], ts), ts))
    for fn in ("timestampdiffSecond", "timestampdiffMinute", "timestampdiffHour", "timestampdiffDay", "timestampdiffWeek",
               "timestampdiffMonth", "timestampdiffQuarter", "timestampdiffYear"):
        outs.append((b.make_function(fn, [t, u], pa.int32()), pa.int32()))
    outs.append((b.make_function("months_between", [t, u], pa.float64()), pa.float64()))
    outs.append((b.make_function("months_between", [d, b.make_function("castDATE", [t], d64)], pa.float64()), pa.float64()))
    return schema, outs, "project"


def case_date_arith_swapped(b):
    """timestampadd* in the (timestamp, count) argument order and with int64 counts (same batch as case_date_arith)."""
    ts, d64 = pa.timestamp("ms"), pa.date64()
    schema = pa.schema([("t", ts), ("u", ts), ("d", d64), ("n", pa.int32()), ("m", pa.int64())])
    t, n, m = F(b, "t", ts), F(b, "n", pa.int32()), F(b, "m", pa.int64())
    n64 = b.make_function("castBIGINT", [n], pa.int64())
    outs = [(b.make_function("timestampaddHour", [t, n], ts), ts),
            (b.make_function("timestampaddWeek", [t, m], ts), ts),
            (b.make_function("timestampaddMonth", [t, n], ts), ts),
            (b.make_function("timestampaddYear", [t, n64], ts), ts),       # small counts: the year must stay in range
            (b.make_function("timestampaddQuarter", [n64, t], ts), ts),
            (b.make_function("add_months", [t, n], ts), ts), (b.make_function("add_months", [F(b, "d", d64), n64], d64), d64)]
    return schema, outs, "project"


def date_arith_batch(n: int, seed: int, null_prob: float = 0.1) -> pa.RecordBatch:
    """timestamps around 1600..2400, month-end days included, small signed counts."""
    rng = np.random.default_rng(seed)
    ts, d64 = pa.timestamp("ms"), pa.date64()
    mk = lambda: rng.integers(-11_676_096_000_000, 13_569_465_600_000, n).astype(np.int64)
    t, u = mk(), mk()
    ends = np.array([951782400000, 1706659200000, 1709164800000, 1711843200000, -2203891200000], dtype=np.int64)
    t[: len(ends)] = ends + rng.integers(0, 86400000, len(ends))   # 2000-02-29, 2024-01-31, 2024-02-29, 2024-03-31, 1900-03-01
    dd = (rng.integers(-150000, 150000, n).astype(np.int64)) * 86400000
    cnt = rng.integers(-400, 400, n).astype(np.int32)
    big = rng.integers(-40000, 40000, n).astype(np.int64)
    mask = lambda: (rng.random(n) < null_prob) if null_prob > 0 else None
    cols = [pa.array(t, pa.int64(), mask=mask()).cast(ts), pa.array(u, pa.int64(), mask=mask()).cast(ts),
            pa.array(dd, pa.int64(), mask=mask()).cast(d64), pa.array(cnt, mask=mask()), pa.array(big, mask=mask())]
    return pa.RecordBatch.from_arrays(cols, names=["t", "u", "d", "n", "m"])


def case_decimal(p1, s1, p2, s2, op, rp, rs):
    def build(b):
        t1, t2, rt = pa.decimal128(p1, s1), pa.decimal128(p2, s2), pa.decimal128(rp, rs)
        schema = pa.schema([("x", t1), ("y", t2)])
        return schema, [(b.make_function(op, [F(b, "x", t1), F(b, "y", t2)], rt), rt)], "project"
    build.__name__ = "decimal_%s_%d_%d_%d_%d" % (op, p1, s1, p2, s2)
    return build


def case_decimal_misc(b):
    t1, t2 = pa.decimal128(15, 2), pa.decimal128(20, 6)
    schema = pa.schema([("x", t1), ("y", t2), ("l", pa.int64())])
    x, y, l = F(b, "x", t1), F(b, "y", t2), F(b, "l", pa.int64())
    B = pa.bool_()
    one = b.make_literal(decimal.Decimal("1.00"), t1)
    return schema, [(b.make_function("less_than", [x, y], B), B),
                    (b.make_function("equal", [x, x], B), B),
                    (b.make_function("greater_than_or_equal_to", [y, x], B), B),
                    (b.make_function("castDECIMAL", [x], pa.decimal128(10, 0)), pa.decimal128(10, 0)),
                    (b.make_function("castDECIMAL", [x], pa.decimal128(30, 8)), pa.decimal128(30, 8)),
                    (b.make_function("castDECIMAL", [l], pa.decimal128(38, 4)), pa.decimal128(38, 4)),
                    (b.make_function("castBIGINT", [y], pa.int64()), pa.int64()),
                    (b.make_function("castFLOAT8", [x], pa.float64()), pa.float64()),
                    (b.make_function("subtract", [one, x], pa.decimal128(16, 2)), pa.decimal128(16, 2)),
                    (b.make_function("abs", [x], t1), t1), (b.make_function("negative", [y], t2), t2)], "project"


def case_decimal_rounding(b):
    """round / truncate / ceil / floor of decimal128: to scale 0, to a literal scale (positive, zero,
    negative, larger than the input's), with result types as the caller declares them."""
    t1, t2 = pa.decimal128(15, 4), pa.decimal128(38, 10)
    schema = pa.schema([("x", t1), ("y", t2)])
    x, y = F(b, "x", t1), F(b, "y", t2)
    fn = b.make_function
    I = pa.int32()
    D = pa.decimal128
    outs = [
        (fn("round", [x], D(12, 0)), D(12, 0)), (fn("truncate", [x], D(12, 0)), D(12, 0)),
        (fn("ceil", [x], D(12, 0)), D(12, 0)), (fn("floor", [x], D(12, 0)), D(12, 0)),
        (fn("round", [x, b.make_literal(2, I)], D(15, 2)), D(15, 2)), (fn("round", [x, b.make_literal(-2, I)], D(12, 0)), D(12, 0)),
        (fn("truncate", [x, b.make_literal(1, I)], D(13, 1)), D(13, 1)), (fn("trunc", [x, b.make_literal(-3, I)], D(12, 0)), D(12, 0)),
        (fn("round", [x, b.make_literal(6, I)], D(17, 6)), D(17, 6)),
        (fn("round", [y], D(29, 0)), D(29, 0)), (fn("ceil", [y], D(29, 0)), D(29, 0)), (fn("floor", [y], D(29, 0)), D(29, 0)),
        (fn("round", [y, b.make_literal(3, I)], D(32, 3)), D(32, 3)), (fn("truncate", [y, b.make_literal(-5, I)], D(29, 0)), D(29, 0)),
        (fn("round", [y, b.make_literal(-40, I)], D(29, 0)), D(29, 0)),
    ]
    return schema, outs, "project"


def decimal_divide_type(p1, s1, p2, s2):
    """The reference's result type for decimal divide (DESIGN.md semantics table):
    scale = max(6, s1 + p2 + 1), precision = p1 - s1 + s2 + scale, capped at 38 by giving up
    scale down to min(scale, 6)."""
    scale = max(6, s1 + p2 + 1)
    prec = p1 - s1 + s2 + scale
    if prec > 38:
        delta = prec - 38
        scale = max(scale - delta, min(scale, 6))
        prec = 38
    return prec, scale


def decimal_mod_type(p1, s1, p2, s2):
    scale = max(s1, s2)
    return min(p1 - s1, p2 - s2) + scale, scale


def case_decimal_divide(p1, s1, p2, s2, guarded=True):
    """x / y (decimal128); guarded: `if y != 0 then x / y else x_cast` so random zeros do not raise."""
    def build(b):
        t1, t2 = pa.decimal128(p1, s1), pa.decimal128(p2, s2)
        rt = pa.decimal128(*decimal_divide_type(p1, s1, p2, s2))
        schema = pa.schema([("x", t1), ("y", t2)])
        x, y = F(b, "x", t1), F(b, "y", t2)
        div = b.make_function("divide", [x, y], rt)
        if not guarded:
            return schema, [(div, rt)], "project"
        nz = b.make_function("not_equal", [y, b.make_literal(decimal.Decimal(0).scaleb(-s2), t2)], pa.bool_())
        return schema, [(b.make_if(nz, div, b.make_literal(None, rt), rt), rt)], "project"
    build.__name__ = "decimal_divide_%d_%d_%d_%d" % (p1, s1, p2, s2)
    return build


def case_decimal_mod(p1, s1, p2, s2):
    def build(b):
        t1, t2 = pa.decimal128(p1, s1), pa.decimal128(p2, s2)
        rt = pa.decimal128(*decimal_mod_type(p1, s1, p2, s2))
        schema = pa.schema([("x", t1), ("y", t2)])
        x, y = F(b, "x", t1), F(b, "y", t2)
        nz = b.make_function("not_equal", [y, b.make_literal(decimal.Decimal(0).scaleb(-s2), t2)], pa.bool_())
        return schema, [(b.make_if(nz, b.make_function("mod", [x, y], rt), b.make_literal(None, rt), rt), rt)], "project"
    build.__name__ = "decimal_mod_%d_%d_%d_%d" % (p1, s1, p2, s2)
    return build


def case_decimal_from_double(b):
    schema = pa.schema([("d", pa.float64()), ("f", pa.float32())])
    d, f = F(b, "d", pa.float64()), F(b, "f", pa.float32())
    outs = []
    for p, s in [(38, 6), (20, 2), (10, 0), (38, 30), (9, 4)]:
        t = pa.decimal128(p, s)
        outs.append((b.make_function("castDECIMAL", [d], t), t))
    t = pa.decimal128(30, 8)
    outs.append((b.make_function("castDECIMAL", [f], t), t))
    return schema, outs, "project"


HASH_TYPES = [pa.int8(), pa.int32(), pa.int64(), pa.uint16(), pa.uint64(), pa.float32(), pa.float64(),
              pa.bool_(), pa.date32(), pa.date64(), pa.timestamp("ms"), pa.string(), pa.binary()]


def case_hash(t):
    def build(b):
        schema = pa.schema([("v", t), ("s32", pa.int32()), ("s64", pa.int64())])
        v, s32, s64 = F(b, "v", t), F(b, "s32", pa.int32()), F(b, "s64", pa.int64())
        I, L = pa.int32(), pa.int64()
        outs = [(b.make_function("hash32", [v], I), I), (b.make_function("hash", [v], I), I),
                (b.make_function("hash64", [v], L), L),
                (b.make_function("hash32", [v, s32], I), I), (b.make_function("hash64", [v, s64], L), L),
                (b.make_function("hash32", [v, b.make_literal(7, I)], I), I),
                (b.make_function("hash64", [v, b.make_literal(-3, L)], L), L)]
        if pa.types.is_string(t):
            outs.append((b.make_function("hash64", [b.make_function("upper", [v], t)], L), L))
            outs.append((b.make_function("hash32", [b.make_function("substr", [v, b.make_literal(2, L), b.make_literal(9, L)], t)], I), I))
        return schema, outs, "project"
    build.__name__ = "hash_%s" % t
    return build


def case_cast_varchar(b):
    t = pa.string()
    schema = pa.schema([("s", t), ("k", pa.int64())])
    s, k = F(b, "s", t), F(b, "k", pa.int64())
    I = pa.int32()
    L = lambda v: b.make_literal(v, pa.int64())
    cv = lambda n: b.make_function("castVARCHAR", [s, n], t)
    # a negative length raises (test_raising_arguments): the variable length is clamped at 0 here
    kpos = b.make_if(b.make_function("greater_than_or_equal_to", [k, L(0)], pa.bool_()), k, L(0), pa.int64())
    return schema, [(b.make_function("octet_length", [cv(L(5))], I), I),
                    (b.make_function("char_length", [cv(kpos)], I), I),
                    (b.make_function("like", [cv(L(7)), b.make_literal("%spa%", t)], pa.bool_()), pa.bool_()),
                    (b.make_function("equal", [cv(L(0)), b.make_literal("", t)], pa.bool_()), pa.bool_())], "project"


def case_in_int(t, values):
    def build(b):
        schema = pa.schema([("a", t)])
        return schema, [(b.make_in_expression(F(b, "a", t), values, t), pa.bool_())], "project"
    build.__name__ = "in_%s_%d" % (t, len(values))
    return build


def case_in_string(b):
    t = pa.string()
    schema = pa.schema([("s", t)])
    return schema, [(b.make_in_expression(F(b, "s", t), ["spark", "fox", "", "日本語 a"], t), pa.bool_())], "project"


def case_like(pattern, escape=None):
    def build(b):
        t = pa.string()
        schema = pa.schema([("s", t)])
        args = [F(b, "s", t), b.make_literal(pattern, t)]
        if escape is not None:
            args.append(b.make_literal(escape, t))
        return schema, [(b.make_function("like", args, pa.bool_()), pa.bool_())], "project"
    build.__name__ = "like_%s" % pattern
    return build


def case_strings(b):
    t = pa.string()
    schema = pa.schema([("s", t), ("u", t), ("k", pa.int64())])
    s, u, k = F(b, "s", t), F(b, "u", t), F(b, "k", pa.int64())
    B, I = pa.bool_(), pa.int32()
    L = lambda v: b.make_literal(v, pa.int64())
    sub = lambda *a: b.make_function("substr", list(a), t)
    outs = [
        (b.make_function("char_length", [s], I), I),
        (b.make_function("octet_length", [s], I), I),
        (b.make_function("starts_with", [s, b.make_literal("sp", t)], B), B),
        (b.make_function("ends_with", [s, b.make_literal("s", t)], B), B),
        (b.make_function("is_substr", [s, b.make_literal("ar", t)], B), B),
        (b.make_function("equal", [s, u], B), B),
        (b.make_function("less_than", [s, u], B), B),
        (b.make_function("greater_than_or_equal_to", [s, u], B), B),
        (b.make_function("char_length", [sub(s, L(2), L(5))], I), I),
        (b.make_function("char_length", [sub(s, L(-3), L(2))], I), I),
        (b.make_function("octet_length", [sub(s, k, L(4))], I), I),
        (b.make_function("equal", [b.make_function("upper", [s], t), b.make_function("upper", [u], t)], B), B),
        (b.make_function("like", [b.make_function("upper", [sub(s, L(1), L(32))], t),
                                  b.make_literal("%SPECIAL%REQUESTS%", t)], B), B),
        (b.make_function("starts_with", [b.make_function("lower", [s], t), b.make_literal("quick", t)], B), B),
        (b.make_function("octet_length", [b.make_function("btrim", [s], t)], I), I),
        (b.make_function("equal", [b.make_function("ltrim", [s], t), b.make_function("rtrim", [s], t)], B), B),
    ]
    return schema, outs, "project"


# ---- LIKE through the warp-cooperative scan (DESIGN.md "String path") --------------------------
LIKE_SCAN_PATTERNS = ["%special%requests%", "%spark%", "spa%ark%fire", "%park%park%", "%requests%special%",
                      "%日本語%", "%fire%fox", "x_y%100%%%park%"]
LIKE_SCAN_VIEWS = ["plain", "upper_substr32", "lower", "substr_3_20", "btrim", "upper"]


def like_scan_view(b, s, view):
    t = pa.string()
    L = lambda v: b.make_literal(v, pa.int64())
    if view == "plain":
        return s
    if view == "upper_substr32":
        return b.make_function("upper", [b.make_function("substr", [s, L(1), L(32)], t)], t)
    if view == "lower":
        return b.make_function("lower", [s], t)
    if view == "upper":
        return b.make_function("upper", [s], t)
    if view == "substr_3_20":
        return b.make_function("substr", [s, L(3), L(20)], t)
    if view == "btrim":
        return b.make_function("btrim", [s], t)
    raise KeyError(view)


def case_like_scan(pattern, view, kind="project"):
    """like(<view>(s), pattern): patterns whose '%'-delimited middle segments are >= 3 bytes go
    through the cooperative scan; the pattern is upper/lower-cased to match the view's case map."""
    def build(b):
        t = pa.string()
        schema = pa.schema([("s", t)])
        pat = pattern.upper() if view in ("upper", "upper_substr32") else pattern
        if pattern.startswith("x_y"):
            args = [like_scan_view(b, F(b, "s", t), view), b.make_literal(pat.replace("_", "\\_"), t),
                    b.make_literal("\\", t)]
        else:
            args = [like_scan_view(b, F(b, "s", t), view), b.make_literal(pat, t)]
        return schema, [(b.make_function("like", args, pa.bool_()), pa.bool_())], kind
    build.__name__ = "like_scan_%s_%s_%s" % (pattern, view, kind)
    return build


def like_scan_batch(n: int, seed: int, null_prob: float = 0.05, offset: int = 0, dense: bool = False,
                    long_rows: bool = False) -> pa.RecordBatch:
    """Strings for the cooperative scan: WORDS mixtures (ASCII and not), optionally rows that are
    nothing but matches (hit-list overflow -> per-lane fallback) and rows longer than the stage."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n + offset):
        if null_prob > 0 and rng.random() < null_prob:
            out.append(None)
            continue
        if dense and i % 97 < 40:
            out.append("special requests park spark fire " * int(rng.integers(1, 3)))
            continue
        if long_rows and i % 251 == 17:
            out.append(("blithely ironic deposits " * 150) + "special final requests park fire")
            continue
        k = int(rng.integers(0, 7))
        out.append(" ".join(WORDS[int(j)] for j in rng.integers(0, len(WORDS), k)))
    arr = pa.array(out, type=pa.string())
    if offset:
        arr = arr.slice(offset)
    return pa.RecordBatch.from_arrays([arr], schema=pa.schema([("s", pa.string())]))


def all_like_scan_cases():
    out = []
    for pat in LIKE_SCAN_PATTERNS:
        for view in LIKE_SCAN_VIEWS:
            if view in ("upper", "upper_substr32") and pat == "%日本語%":
                continue
            out.append(case_like_scan(pat, view))
    return out


def case_string_outputs(b):
    """utf8 outputs (two-pass string projection): views, case maps, if/else, literals, plus a
    fixed-width output from the same projector."""
    t = pa.string()
    schema = pa.schema([("s", t), ("u", t), ("k", pa.int64()), ("a", pa.int32())])
    s, u, k, a = F(b, "s", t), F(b, "u", t), F(b, "k", pa.int64()), F(b, "a", pa.int32())
    L = lambda v: b.make_literal(v, pa.int64())
    B = pa.bool_()
    up = b.make_function("upper", [s], t)
    cond = b.make_function("greater_than", [a, b.make_literal(0, pa.int32())], B)
    outs = [
        (s, t),
        (up, t),
        (b.make_function("lower", [b.make_function("substr", [s, L(2), L(9)], t)], t), t),
        (b.make_function("btrim", [u], t), t),
        (b.make_if(cond, up, b.make_function("castVARCHAR", [u, b.make_if(b.make_function("greater_than_or_equal_to", [k, L(0)], B), k, L(2), pa.int64())], t), t), t),
        (b.make_if(b.make_function("like", [s, b.make_literal("%spark%", t)], B),
                   b.make_literal("SPARK!", t), b.make_if(cond, b.make_literal("", t), b.make_literal(None, t), t), t), t),
        (b.make_function("char_length", [s], pa.int32()), pa.int32()),
        (b.make_function("substr", [u, k], t), t),
    ]
    return schema, outs, "project"


def case_concat_outputs(b):
    """concat / concatOperator results (ropes of views) projected as utf8 outputs."""
    t = pa.string()
    schema = pa.schema([("s", t), ("u", t), ("a", pa.int32())])
    s, u, a = F(b, "s", t), F(b, "u", t), F(b, "a", pa.int32())
    L = lambda v: b.make_literal(v, pa.int64())
    lit = lambda v: b.make_literal(v, t)
    cond = b.make_function("greater_than", [a, b.make_literal(0, pa.int32())], pa.bool_())
    cc = lambda *x: b.make_function("concat", list(x), t)
    co = lambda *x: b.make_function("concatOperator", list(x), t)
    up = b.make_function("upper", [s], t)
    outs = [
        (cc(s, u), t),
        (co(s, lit(" | "), u), t),
        (cc(up, lit("-"), b.make_function("lower", [b.make_function("substr", [u, L(1), L(3)], t)], t), lit("!")), t),
        (cc(cc(s, lit("/")), co(u, lit("/"), s)), t),
        (b.make_if(cond, cc(s, lit("+"), u), up, t), t),
        (b.make_if(cond, lit("x"), co(u, s), t), t),
    ]
    return schema, outs, "project"


def case_binary_output(b):
    t = pa.binary()
    schema = pa.schema([("x", t), ("a", pa.int32())])
    x, a = F(b, "x", t), F(b, "a", pa.int32())
    cond = b.make_function("less_than", [a, b.make_literal(10, pa.int32())], pa.bool_())
    return schema, [(b.make_if(cond, x, b.make_literal(b"\x00\xffraw", t), t), t)], "project"


def case_literals_only(b):
    t = pa.int32()
    schema = pa.schema([("a", t)])
    root = b.make_function("add", [b.make_literal(40, t), b.make_literal(2, t)], t)
    return schema, [(root, t)], "project"


def case_bool_io(b):
    t = pa.bool_()
    schema = pa.schema([("x", t), ("a", pa.int32())])
    x, a = F(b, "x", t), F(b, "a", pa.int32())
    return schema, [(b.make_function("not", [x], t), t),
                    (b.make_if(x, a, b.make_function("negative", [a], pa.int32()), pa.int32()), pa.int32()),
                    (b.make_function("equal", [x, b.make_function("greater_than", [a, b.make_literal(0, pa.int32())], t)], t), t)], "project"


def q6_condition(b, d32=True):
    """TPC-H Q6 predicate over (l_shipdate date32, l_discount f64, l_quantity f64)."""
    sd, f = pa.date32(), pa.float64()
    ship, disc, qty = F(b, "l_shipdate", sd), F(b, "l_discount", f), F(b, "l_quantity", f)
    B = pa.bool_()
    lit_d = lambda days: b.make_literal(days, sd)
    lit_f = lambda v: b.make_literal(v, f)
    conds = [
        b.make_function("greater_than_or_equal_to", [ship, lit_d(8766)], B),   # 1994-01-01
        b.make_function("less_than", [ship, lit_d(9131)], B),                  # 1995-01-01
        b.make_function("greater_than_or_equal_to", [disc, lit_f(0.05)], B),
        b.make_function("less_than_or_equal_to", [disc, lit_f(0.07)], B),
        b.make_function("less_than", [qty, lit_f(24.0)], B),
    ]
    return b.make_and(conds)


Q6_SCHEMA = pa.schema([("l_shipdate", pa.date32()), ("l_discount", pa.float64()),
                       ("l_quantity", pa.float64())])


def case_q6_filter(b):
    return Q6_SCHEMA, [(q6_condition(b), pa.bool_())], "filter"


def case_filter_float(b):
    t = pa.float64()
    schema = pa.schema([("a", t), ("b", t)])
    a, bb = F(b, "a", t), F(b, "b", t)
    B = pa.bool_()
    c1 = b.make_function("less_than", [a, b.make_literal(50.0, t)], B)
    c2 = b.make_function("greater_than", [a, bb], B)
    c3 = b.make_function("less_than", [bb, b.make_literal(11.0, t)], B)
    return schema, [(b.make_or([b.make_and([c1, c2]), c3]), B)], "filter"


def case_filter_all(b):
    t = pa.int32()
    schema = pa.schema([("a", t)])
    return schema, [(b.make_function("isnotnull", [F(b, "a", t)], pa.bool_()), pa.bool_())], "filter"


def case_filter_none(b):
    t = pa.int32()
    schema = pa.schema([("a", t)])
    a = F(b, "a", t)
    return schema, [(b.make_and([b.make_function("isnull", [a], pa.bool_()),
                                 b.make_function("isnotnull", [a], pa.bool_())]), pa.bool_())], "filter"


def case_filter_string(b):
    t = pa.string()
    schema = pa.schema([("c", t)])
    c = F(b, "c", t)
    sub = b.make_function("substr", [c, b.make_literal(1, pa.int64()), b.make_literal(32, pa.int64())], t)
    return schema, [(b.make_function("like", [b.make_function("upper", [sub], t),
                                              b.make_literal("%SPECIAL%REQUESTS%", t)], pa.bool_()), pa.bool_())], "filter"


def case_divide(t):
    def build(b):
        schema = pa.schema([("a", t), ("b", t)])
        a, bb = F(b, "a", t), F(b, "b", t)
        B = pa.bool_()
        # guard: if b != 0 then a / b else a   (the untaken branch must not raise)
        nz = b.make_function("not_equal", [bb, b.make_literal(0 if t not in FLOAT_TYPES else 0.0, t)], B)
        return schema, [(b.make_if(nz, b.make_function("divide", [a, bb], t), a, t), t)], "project"
    build.__name__ = "divide_guarded_%s" % t
    return build


def case_and_short_circuit(b):
    t = pa.int32()
    schema = pa.schema([("a", t), ("b", t)])
    a, bb = F(b, "a", t), F(b, "b", t)
    B = pa.bool_()
    nz = b.make_function("not_equal", [bb, b.make_literal(0, t)], B)
    q = b.make_function("greater_than", [b.make_function("divide", [a, bb], t), b.make_literal(1, t)], B)
    return schema, [(b.make_and([nz, q]), B)], "project"


Q1_SCHEMA = pa.schema([("l_quantity", pa.int64()), ("l_extendedprice", pa.decimal128(15, 2)),
                       ("l_discount", pa.decimal128(15, 2)), ("l_tax", pa.decimal128(15, 2)),
                       ("l_extendedprice_f", pa.float64()), ("l_discount_f", pa.float64()),
                       ("l_tax_f", pa.float64()), ("l_shipdate", pa.date32())])
# generator column kinds (csrc/device/static_kernels.cu) of the Q1 schema, in field order
Q1_KINDS = [3, 4, 5, 6, 7, 1, 8, 0]


def q1_outputs(b):
    """TPC-H Q1-style eight-output projector: decimal and float
    `ext*(1-disc)` and `ext*(1-disc)*(1+tax)`, `qty+qty`, and three CASE expressions.
    Decimal result types follow the reference's rule (p1+p2+1, s1+s2; capped at 38, min scale 6)."""
    D, F64, I64, SD = pa.decimal128(15, 2), pa.float64(), pa.int64(), pa.date32()
    f = {x.name: b.make_field(x) for x in Q1_SCHEMA}
    one_d = b.make_literal(decimal.Decimal("1.00"), D)
    one_f = b.make_literal(1.0, F64)
    B = pa.bool_()
    d1 = b.make_function("multiply", [f["l_extendedprice"],
                                      b.make_function("subtract", [one_d, f["l_discount"]], pa.decimal128(16, 2))],
                         pa.decimal128(32, 4))
    d2 = b.make_function("multiply", [d1, b.make_function("add", [one_d, f["l_tax"]], pa.decimal128(16, 2))],
                         pa.decimal128(38, 6))
    f1 = b.make_function("multiply", [f["l_extendedprice_f"],
                                      b.make_function("subtract", [one_f, f["l_discount_f"]], F64)], F64)
    f2 = b.make_function("multiply", [f1, b.make_function("add", [one_f, f["l_tax_f"]], F64)], F64)
    q2 = b.make_function("add", [f["l_quantity"], f["l_quantity"]], I64)
    c1 = b.make_if(b.make_function("greater_than", [f["l_discount_f"], b.make_literal(0.05, F64)], B),
                   f["l_extendedprice_f"], b.make_literal(0.0, F64), F64)
    c2 = b.make_if(b.make_function("less_than", [f["l_quantity"], b.make_literal(24, I64)], B),
                   b.make_literal(1, I64), b.make_literal(0, I64), I64)
    c3 = b.make_if(b.make_function("less_than_or_equal_to", [f["l_shipdate"], b.make_literal(10471, SD)], B),
                   f["l_quantity"], b.make_literal(None, I64), I64)
    return [(d1, pa.decimal128(32, 4)), (d2, pa.decimal128(38, 6)), (f1, F64), (f2, F64), (q2, I64),
            (c1, F64), (c2, I64), (c3, I64)]


def case_q1_projector(b):
    return Q1_SCHEMA, q1_outputs(b), "project"


def q1_batch(n: int, seed: int = 42, null_permille: int = 20) -> pa.RecordBatch:
    """Synthetic lineitem columns of the Q1 schema from the CPU generator (oracle/lineitem.h)."""
    import oracle as o
    cols = []
    for kind, f in zip(Q1_KINDS, Q1_SCHEMA):
        vals, vld = o.generate_lineitem(kind, seed, 0, n, null_permille, threads=4)
        bufs = [pa.py_buffer(vld) if vld is not None else None, pa.py_buffer(vals)]
        cols.append(pa.Array.from_buffers(f.type, n, bufs))
    return pa.RecordBatch.from_arrays(cols, schema=Q1_SCHEMA)


COMMENT_SCHEMA = pa.schema([("l_comment", pa.string())])
_COMMENT_WORDS = ["furiously", "carefully", "quickly", "blithely", "slyly", "ironic", "final", "regular",
                  "express", "bold", "pending", "even", "deposits", "accounts", "packages", "theodolites",
                  "instructions", "foxes", "pinto", "beans", "platelets", "asymptotes", "dependencies",
                  "sleep", "nag", "haggle", "wake", "cajole", "above", "the", "among", "about"]


def comment_batch(n: int, seed: int = 42, null_prob: float = 0.01) -> pa.RecordBatch:
    """l_comment-like utf8 column: ASCII words, length 10..43 bytes
    (mean ~27), ~1 % of the rows contain "special" ... "requests", ~1 % nulls."""
    rng = np.random.default_rng(seed)
    pool = []
    for i in range(4096):
        target = int(rng.integers(10, 44))
        words = []
        if i % 100 == 7:
            words = ["special", _COMMENT_WORDS[int(rng.integers(0, len(_COMMENT_WORDS)))], "requests"]
        while len(" ".join(words)) < target:
            words.append(_COMMENT_WORDS[int(rng.integers(0, len(_COMMENT_WORDS)))])
        s = " ".join(words)[:43] if i % 100 != 7 else " ".join(words)[:43]
        pool.append(s if len(s) >= 10 else s + " sleep nag")
    pool_arr = pa.array(pool, type=pa.string())
    idx = rng.integers(0, len(pool), n)
    arr = pool_arr.take(pa.array(idx))
    if null_prob > 0:
        mask = rng.random(n) < null_prob
        arr = pa.Array.from_buffers(pa.string(), n, [pa.py_buffer(np.packbits(~mask, bitorder="little")),
                                                     arr.buffers()[1], arr.buffers()[2]])
    return pa.RecordBatch.from_arrays([arr], schema=COMMENT_SCHEMA)


def comment_condition(b):
    """like(upper(substr(l_comment, 1, 32)), '%SPECIAL%REQUESTS%')"""
    t = pa.string()
    c = b.make_field(COMMENT_SCHEMA.field(0))
    sub = b.make_function("substr", [c, b.make_literal(1, pa.int64()), b.make_literal(32, pa.int64())], t)
    return b.make_function("like", [b.make_function("upper", [sub], t),
                                    b.make_literal("%SPECIAL%REQUESTS%", t)], pa.bool_())


def all_project_cases():
    cases = []
    for t in NUMERIC:
        for op in ("add", "subtract", "multiply"):
            cases.append(case_arith(op, t))
    for t in [pa.int32(), pa.uint64(), pa.float32(), pa.float64(), pa.int8()]:
        for op in RELOPS:
            cases.append(case_relop(op, t))
    for t in [pa.date32(), pa.date64(), pa.timestamp("ms")]:
        cases.append(case_relop("less_than", t))
        cases.append(case_relop("equal", t))
    cases += [case_if_else, case_if_null_literal, case_kleene, case_null_tests, case_casts, case_mod,
              case_dates, case_decimal_misc, case_in_string, case_strings, case_literals_only,
              case_bool_io, case_and_short_circuit, case_q1_projector]
    for t in [pa.int8(), pa.int32(), pa.int64(), pa.uint32(), pa.float32(), pa.float64()]:
        cases.append(case_divide(t))
    # decimal result types follow the reference's rule: add/sub: s=max(s1,s2),
    # p=max(p1-s1,p2-s2)+s+1; multiply: p=p1+p2+1, s=s1+s2; capped at 38 with min scale 6
    cases += [
        case_decimal(12, 2, 12, 2, "multiply", 25, 4),
        case_decimal(15, 2, 15, 2, "multiply", 31, 4),
        case_decimal(31, 4, 16, 2, "multiply", 38, 6),
        case_decimal(38, 10, 38, 10, "multiply", 38, 6),
        case_decimal(30, 12, 20, 9, "multiply", 38, 8),
        case_decimal(15, 2, 15, 2, "add", 16, 2),
        case_decimal(15, 2, 20, 6, "add", 21, 6),
        case_decimal(38, 10, 38, 4, "subtract", 38, 6),
        case_decimal(38, 0, 38, 0, "add", 38, 0),
        case_decimal(10, 5, 12, 1, "subtract", 17, 5),
    ]
    cases += [case_decimal_divide(15, 2, 15, 2), case_decimal_divide(38, 10, 20, 4),
              case_decimal_divide(10, 0, 5, 3), case_decimal_divide(30, 20, 38, 2),
              case_decimal_divide(38, 30, 12, 0),
              case_decimal_mod(15, 2, 15, 2), case_decimal_mod(38, 10, 20, 4), case_decimal_mod(20, 0, 38, 30),
              case_decimal_from_double, case_cast_varchar, case_string_outputs, case_binary_output,
              case_concat_outputs, case_rounding, case_date_arith, case_date_arith_swapped, case_intmath, case_calendar,
              case_string_positions, case_number_to_text, case_string_misc, case_virtual_strings,
              case_decimal_rounding, case_math, case_trig, case_regexp, case_misc_casts, case_power, case_inverse_trig, case_in_floats, case_concat_consumers]
    cases += [case_hash(t) for t in HASH_TYPES]
    cases += [case_in_int(pa.int32(), [1, 5]), case_in_int(pa.int64(), [1, 5, -3]),
              case_in_int(pa.int32(), list(range(-20, 40, 3)))]
    for pat in ["%spark%", "spark%", "%spark", "s_ark%", "%", "", "_", "%a%b%c%", "fire", "%日本%",
                "_本%", "%%x%%", "a%a", "%ss"]:
        cases.append(case_like(pat))
    cases.append(case_like("100\\%%", "\\"))
    cases.append(case_like("x\\_y%", "\\"))
    return cases


def all_filter_cases():
    return [case_q6_filter, case_filter_float, case_filter_all, case_filter_none, case_filter_string]


def q6_batch(n: int, seed: int = 42, null_permille: int = 0, oracle_mod=None) -> pa.RecordBatch:
    """Synthetic lineitem from the CPU generator in oracle/."""
    import oracle as _o
    o = oracle_mod or _o
    cols = []
    for kind, t in ((0, pa.date32()), (1, pa.float64()), (2, pa.float64())):
        vals, vld = o.generate_lineitem(kind, seed, 0, n, null_permille, threads=4)
        bufs = [pa.py_buffer(vld) if vld is not None else None, pa.py_buffer(vals)]
        cols.append(pa.Array.from_buffers(t, n, bufs))
    return pa.RecordBatch.from_arrays(cols, schema=Q6_SCHEMA)


REGEX_PATTERNS = [
    "abc", "a.c", "^abc", "abc$", "^abc$", "a*", "^a*$", "a+b", "ab?c", "a|b", "^(a|bc)+$", "(ab)*c", "[abc]+", "[^abc]", "^[^abc]*$",
    "[a-c0-2]{2,3}", "a{3}", "a{2,}", "^.{3}$", ".*", "^.*$", "", "^$", "\\d+", "\\w+\\s\\w+", "\\D", "^\\S+$", "\\.", "a\\|b", "[.]", "[]a]", "[a\\]]",
    "[\\d_]+", "é", "^é+$", "日.語", "[é日]", "[^a]é", "a.*b.*c", "(a|b)*abb", "(?:ab|cd)+e", "x?y?z?$", "^(a?){3}b", "(a*)*b", "(a|ab)(c|bcd)(d*)",
    "[a-c]+[0-2]*$", "\\n", "a\\nb", "^[^\\n]*$", "\\t", "-", "[a-]", "[-a]", "_+", "a{1,2}b{0,1}", "((a))", "(a|)", "(|a)b", "0[01]*1", "^\\w*$",
    "(ab|a)(bc|c)?$", "^(?:[a-c]|é)+$", ".\\W.", "\\s$", "^\\s", "c.{2}c", "[0-9]+(\\.[0-9]+)?$", "^-?\\d+$",
]


def _random_regex(rng, depth=0):
    """A random pattern of the supported subset over the alphabet regex_texts() draws from."""
    atoms = ["a", "b", "c", "0", "1", " ", "_", "-", "\\.", ".", "é", "日", "x", "[abc]", "[^ab]", "[a-c0-2]", "[^\\n0-9]", "\\d", "\\w", "\\s",
             "\\D", "\\W", "\\S", "[é_-]", "\\|", "\\n", "[\\d\\s]"]
    parts = []
    for _ in range(int(rng.integers(1, 4 if depth else 5))):
        k = int(rng.integers(0, 10))
        if k < 6 or depth >= 2:
            atom = str(rng.choice(atoms))
        elif k < 8:
            atom = "(" + _random_regex(rng, depth + 1) + ")"
        else:
            atom = "(?:" + _random_regex(rng, depth + 1) + "|" + _random_regex(rng, depth + 1) + ")"
        q = int(rng.integers(0, 12))
        atom += ["", "", "", "", "", "*", "+", "?", "{2}", "{1,2}", "{0,1}", "*?"][q]
        parts.append(atom)
    return "".join(parts)


def _random_regexes(n, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        p = _random_regex(rng)
        k = int(rng.integers(0, 8))
        p = ("^" + p if k in (0, 2) else p) + ("$" if k in (1, 2) else "")
        out.append(p)
    return out


REGEX_PATTERNS += _random_regexes(61, 2024)
REGEX_PATTERNS += ["(?i)abc", "(?i)^[a-c]+$", "(?i)é|日本", "(?i)[^a]b", "(?i)A\\wC$", "\\x61\\x62", "[\\x30-\\x32]+$", "[[:alpha:]]+", "^[[:digit:][:space:]]+$",
                   "[^[:alnum:]_]", "[[:upper:]][[:lower:]]", "(?i)[[:upper:]]{2}", "[[:punct:]]{2,}", "^[[:word:]-]*$", "[[:xdigit:]]{3}",
                   # more than 64 automaton positions: the two-word matcher
                   "^(?:[a-c][0-2 .]?_?){1,24}", "^(?:\\w|\\s|é){1,20}$", "(?:a|b|c|0|1|2| |_|-|x|y|z|d|e){3}[|\\]A.]?(?:ab|bc|ca|é日|語x){10}"]


def regex_texts(n, seed):
    """Short texts over a small alphabet (so that the patterns above hit often), some multi-byte, some NULL."""
    import numpy as np
    rng = np.random.default_rng(seed)
    alphabet = list("aaabbbccc012 _-.\n\t") + ["é", "日", "語", "x", "y", "z", "d", "e", "|", "]", "A"]
    out = ["", "abc", "aabbcc", "abb", "日本語", "a\nb", "é", "ééé"]
    while len(out) < n:
        out.append("".join(rng.choice(alphabet, size=int(rng.integers(0, 12)))))
    out = out[:n]
    for k in range(5, n, 13):
        out[k] = None
    return out
