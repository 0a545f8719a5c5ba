"""Second opinion on the oracle: pyarrow.compute (Arrow 24.0.0), an independent CPU
implementation, on the operations where its semantics coincide with the table in DESIGN.md
(SURVEY.md §8c mitigation 1).  Where Arrow compute is known to differ the case is left out
and the difference is noted in DESIGN.md (e.g. plain and_/or_ are not Kleene; decimal result
types above precision 38; division by zero raises in both but Arrow has no per-row guard)."""
import decimal

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

import cases
from helpers import assert_arrays_match

N = 5000


def run_oracle(oracle, gandiva, build, batch):
    b = gandiva.TreeExprBuilder()
    schema, outs, kind = build(b)
    return oracle.project([r for r, _ in outs], [t for _, t in outs], batch)


@pytest.mark.parametrize("t", cases.NUMERIC, ids=str)
@pytest.mark.parametrize("op,fn", [("add", pc.add), ("subtract", pc.subtract), ("multiply", pc.multiply)])
def test_wrapping_arithmetic(op, fn, t, oracle, gandiva):
    build = cases.case_arith(op, t)
    schema = pa.schema([("a", t), ("b", t)])
    batch = cases.random_batch(schema, N, seed=1)
    got, = run_oracle(oracle, gandiva, build, batch)
    want = fn(batch.column(0), batch.column(1))  # unchecked variants wrap around
    assert_arrays_match(got, want.cast(t), "%s %s" % (op, t))


@pytest.mark.parametrize("t", [pa.int32(), pa.int64(), pa.uint8(), pa.float32(), pa.float64(),
                               pa.date32(), pa.timestamp("ms")], ids=str)
@pytest.mark.parametrize("op,fn", [("equal", pc.equal), ("not_equal", pc.not_equal),
                                   ("less_than", pc.less), ("less_than_or_equal_to", pc.less_equal),
                                   ("greater_than", pc.greater),
                                   ("greater_than_or_equal_to", pc.greater_equal)])
def test_comparisons(op, fn, t, oracle, gandiva):
    build = cases.case_relop(op, t)
    schema = pa.schema([("a", t), ("b", t)])
    batch = cases.random_batch(schema, N, seed=2, small=True)
    got, = run_oracle(oracle, gandiva, build, batch)
    assert_arrays_match(got, fn(batch.column(0), batch.column(1)), "%s %s" % (op, t))


def test_kleene_logic(oracle, gandiva):
    schema = pa.schema([("x", pa.bool_()), ("y", pa.bool_()), ("z", pa.bool_())])
    batch = cases.random_batch(schema, N, seed=3, null_prob=0.3)
    got = run_oracle(oracle, gandiva, cases.case_kleene, batch)
    x, y, z = batch.columns
    assert_arrays_match(got[0], pc.and_kleene(x, y), "and")
    assert_arrays_match(got[1], pc.or_kleene(x, y), "or")
    assert_arrays_match(got[2], pc.and_kleene(pc.and_kleene(x, y), z), "and3")
    assert_arrays_match(got[3], pc.or_kleene(pc.and_kleene(x, y), pc.invert(z)), "or(and,not)")


def test_if_else(oracle, gandiva):
    t = pa.int32()
    schema = pa.schema([("a", t), ("b", t), ("c", t)])
    batch = cases.random_batch(schema, N, seed=4, null_prob=0.2, small=True)
    got, = run_oracle(oracle, gandiva, cases.case_if_else, batch)
    a, b, c = batch.columns
    # a null condition selects the else branch: fill_null(False) before if_else
    cond1 = pc.fill_null(pc.greater(a, b), False)
    cond2 = pc.fill_null(pc.less(b, c), False)
    inner = pc.if_else(cond2, b, pc.add(c, pa.scalar(7, t)))
    assert_arrays_match(got, pc.if_else(cond1, a, inner), "if_else")


def test_null_tests(oracle, gandiva):
    schema = pa.schema([("a", pa.float64()), ("b", pa.float64()), ("x", pa.bool_())])
    batch = cases.random_batch(schema, N, seed=5, null_prob=0.3, small=True)
    got = run_oracle(oracle, gandiva, cases.case_null_tests, batch)
    a, b, x = batch.columns
    assert_arrays_match(got[0], pc.is_null(a), "isnull")
    assert_arrays_match(got[1], pc.is_valid(b), "isnotnull")
    assert_arrays_match(got[4], pc.fill_null(x, False), "istrue")
    assert_arrays_match(got[5], pc.fill_null(x, True), "isnotfalse")


def test_casts(oracle, gandiva):
    schema = pa.schema([("i", pa.int32()), ("l", pa.int64()), ("f", pa.float32()), ("d", pa.float64())])
    batch = cases.random_batch(schema, N, seed=6)
    got = run_oracle(oracle, gandiva, cases.case_casts, batch)
    i, l, f, d = batch.columns
    assert_arrays_match(got[0], i.cast(pa.int64()), "castBIGINT")
    assert_arrays_match(got[2], l.cast(pa.float32(), safe=False), "castFLOAT4(int64)")
    assert_arrays_match(got[3], d.cast(pa.float32(), safe=False), "castFLOAT4(double)")
    assert_arrays_match(got[4], l.cast(pa.float64(), safe=False), "castFLOAT8(int64)")
    assert_arrays_match(got[5], f.cast(pa.float64()), "castFLOAT8(float)")
    assert_arrays_match(got[7], pc.negate(d), "negative")


def test_date_extraction(oracle, gandiva):
    ts, d64, d32 = pa.timestamp("ms"), pa.date64(), pa.date32()
    schema = pa.schema([("t", ts), ("d", d64), ("e", d32)])
    batch = cases.random_batch(schema, N, seed=7)
    got = run_oracle(oracle, gandiva, cases.case_dates, batch)
    t, d, e = batch.columns
    L = pa.int64()
    assert_arrays_match(got[0], pc.year(t).cast(L), "year")
    assert_arrays_match(got[1], pc.month(t).cast(L), "month")
    assert_arrays_match(got[2], pc.day(t).cast(L), "day")
    assert_arrays_match(got[3], pc.hour(t).cast(L), "hour")
    assert_arrays_match(got[4], pc.minute(t).cast(L), "minute")
    assert_arrays_match(got[5], pc.second(t).cast(L), "second")
    assert_arrays_match(got[6], pc.day_of_year(t).cast(L), "doy")
    # Arrow: Monday=0 by default; ours Sunday=1
    assert_arrays_match(got[7], pc.add(pc.day_of_week(t, count_from_zero=True, week_start=7), 1).cast(L), "dow")
    assert_arrays_match(got[8], pc.quarter(t).cast(L), "quarter")
    assert_arrays_match(got[10], pc.year(d).cast(L), "year(date64)")
    assert_arrays_match(got[11], pc.month(e).cast(L), "month(date32)")
    assert_arrays_match(got[12], pc.day(e).cast(L), "day(date32)")


@pytest.mark.parametrize("p1,s1,p2,s2", [(12, 2, 12, 2), (15, 2, 15, 2), (10, 3, 8, 0), (18, 6, 17, 9)])
def test_decimal_multiply_no_rescale(p1, s1, p2, s2, oracle, gandiva):
    """p1+p2+1 <= 38: the reference's result type equals Arrow compute's (probed in SURVEY.md
    §8a: decimal128(12,2)^2 -> decimal128(25,4)) and no rounding happens."""
    rp, rs = p1 + p2 + 1, s1 + s2
    build = cases.case_decimal(p1, s1, p2, s2, "multiply", rp, rs)
    schema = pa.schema([("x", pa.decimal128(p1, s1)), ("y", pa.decimal128(p2, s2))])
    batch = cases.random_batch(schema, 2000, seed=8)
    got, = run_oracle(oracle, gandiva, build, batch)
    want = pc.multiply(batch.column(0), batch.column(1))
    assert want.type == pa.decimal128(rp, rs)
    assert_arrays_match(got, want, "decimal multiply")


@pytest.mark.parametrize("op,fn", [("add", pc.add), ("subtract", pc.subtract)])
@pytest.mark.parametrize("p1,s1,p2,s2", [(15, 2, 15, 2), (15, 2, 20, 6), (10, 5, 12, 1)])
def test_decimal_addsub_no_rescale(op, fn, p1, s1, p2, s2, oracle, gandiva):
    s = max(s1, s2)
    p = max(p1 - s1, p2 - s2) + s + 1
    build = cases.case_decimal(p1, s1, p2, s2, op, p, s)
    schema = pa.schema([("x", pa.decimal128(p1, s1)), ("y", pa.decimal128(p2, s2))])
    batch = cases.random_batch(schema, 2000, seed=9)
    got, = run_oracle(oracle, gandiva, build, batch)
    want = fn(batch.column(0), batch.column(1))
    assert want.type == pa.decimal128(p, s)
    assert_arrays_match(got, want, "decimal %s" % op)


def test_decimal_rescale_rounds_half_away(oracle, gandiva):
    """Scale reduction: checked against Python's decimal with ROUND_HALF_UP on the magnitude."""
    build = cases.case_decimal(38, 10, 38, 10, "multiply", 38, 6)
    schema = pa.schema([("x", pa.decimal128(38, 10)), ("y", pa.decimal128(38, 10))])
    batch = cases.random_batch(schema, 2000, seed=10, small=False)
    # keep magnitudes small enough that most products fit 38 digits
    rng = np.random.default_rng(11)
    xs = [decimal.Decimal(int(rng.integers(-10**14, 10**14))).scaleb(-10) if rng.random() > 0.1 else None for _ in range(2000)]
    ys = [decimal.Decimal(int(rng.integers(-10**14, 10**14))).scaleb(-10) if rng.random() > 0.1 else None for _ in range(2000)]
    batch = pa.RecordBatch.from_arrays([pa.array(xs, schema.field(0).type), pa.array(ys, schema.field(1).type)], schema=schema)
    got, = run_oracle(oracle, gandiva, build, batch)
    ctx = decimal.Context(prec=100)
    want = []
    for x, y in zip(xs, ys):
        if x is None or y is None:
            want.append(None)
            continue
        q = ctx.multiply(x, y).quantize(decimal.Decimal(1).scaleb(-6), rounding=decimal.ROUND_HALF_UP, context=ctx)
        want.append(q)
    assert got.to_pylist() == want


def test_like_scan_patterns(oracle, gandiva):
    """The patterns/views of the cooperative-scan GPU tests, oracle against pyarrow.compute."""
    batch = cases.like_scan_batch(3000, seed=11, dense=True, long_rows=True)
    s = batch.column(0)
    for pat in cases.LIKE_SCAN_PATTERNS:
        if pat.startswith("x_y"):
            continue  # match_like has no escape argument
        got, = run_oracle(oracle, gandiva, cases.case_like_scan(pat, "plain"), batch)
        assert_arrays_match(got, pc.match_like(s, pat), "like " + pat)
        got, = run_oracle(oracle, gandiva, cases.case_like_scan(pat, "lower"), batch)
        assert_arrays_match(got, pc.match_like(pc.ascii_lower(s), pat), "like lower " + pat)
        if pat != "%日本語%":
            got, = run_oracle(oracle, gandiva, cases.case_like_scan(pat, "upper_substr32"), batch)
            want = pc.match_like(pc.ascii_upper(pc.utf8_slice_codeunits(s, 0, 32)), pat.upper())
            assert_arrays_match(got, want, "like upper substr " + pat)
        got, = run_oracle(oracle, gandiva, cases.case_like_scan(pat, "btrim"), batch)
        assert_arrays_match(got, pc.match_like(pc.ascii_trim(s, " "), pat), "like btrim " + pat)


def test_like_and_strings(oracle, gandiva):
    t = pa.string()
    schema = pa.schema([("s", t)])
    batch = cases.random_batch(schema, 3000, seed=12)
    s = batch.column(0)
    for pat in ["%spark%", "spark%", "%spark", "s_ark%", "%", "_", "%a%b%c%", "%日本%", "_本%", "a%a"]:
        got, = run_oracle(oracle, gandiva, cases.case_like(pat), batch)
        assert_arrays_match(got, pc.match_like(s, pat), "like " + pat)


def test_string_functions(oracle, gandiva):
    t = pa.string()
    schema = pa.schema([("s", t), ("u", t), ("k", pa.int64())])
    rng = np.random.default_rng(13)
    base = cases.random_batch(pa.schema([("s", t), ("u", t)]), 3000, seed=13)
    k = pa.array(rng.integers(-6, 9, 3000), type=pa.int64())
    batch = pa.RecordBatch.from_arrays([base.column(0), base.column(1), k], schema=schema)
    got = run_oracle(oracle, gandiva, cases.case_strings, batch)
    s, u, _ = batch.columns
    assert_arrays_match(got[0], pc.utf8_length(s), "char_length")
    assert_arrays_match(got[1], pc.binary_length(s), "octet_length")
    assert_arrays_match(got[2], pc.starts_with(s, "sp"), "starts_with")
    assert_arrays_match(got[3], pc.ends_with(s, "s"), "ends_with")
    assert_arrays_match(got[4], pc.match_substring(s, "ar"), "is_substr")
    assert_arrays_match(got[5], pc.equal(s, u), "equal")
    assert_arrays_match(got[6], pc.less(s, u), "less_than")
    assert_arrays_match(got[7], pc.greater_equal(s, u), "greater_equal")
    # substr(s, 2, 5) == codeunit slice [1, 6)
    assert_arrays_match(got[8], pc.utf8_length(pc.utf8_slice_codeunits(s, 1, 6)), "substr(2,5)")
    # upper: ASCII-only here; compare against Arrow's ascii_upper
    assert_arrays_match(got[11], pc.equal(pc.ascii_upper(s), pc.ascii_upper(u)), "upper eq")
    want_like = pc.match_like(pc.ascii_upper(pc.utf8_slice_codeunits(s, 0, 32)), "%SPECIAL%REQUESTS%")
    assert_arrays_match(got[12], want_like, "like(upper(substr))")
    assert_arrays_match(got[14], pc.binary_length(pc.utf8_trim(s, " ")), "btrim")


def test_in_expression(oracle, gandiva):
    t = pa.int32()
    vals = list(range(-20, 40, 3))
    schema = pa.schema([("a", t)])
    batch = cases.random_batch(schema, N, seed=14, small=True)
    got, = run_oracle(oracle, gandiva, cases.case_in_int(t, vals), batch)
    want = pc.is_in(batch.column(0), value_set=pa.array(vals, t))
    # is_in returns false for nulls; the expression is null there
    want = pc.if_else(pc.is_valid(batch.column(0)), want, pa.scalar(None, pa.bool_()))
    assert_arrays_match(got, want, "in")


def test_divide_by_zero_raises(oracle, gandiva):
    b = gandiva.TreeExprBuilder()
    t = pa.int32()
    schema = pa.schema([("a", t), ("b", t)])
    root = b.make_function("divide", [cases.F(b, "a", t), cases.F(b, "b", t)], t)
    batch = pa.RecordBatch.from_arrays([pa.array([1, 2, 3], t), pa.array([1, 0, 2], t)], schema=schema)
    with pytest.raises(Exception, match="divide by zero"):
        oracle.project([root], [t], batch)
    # a null divisor slot holding 0 must not raise
    batch = pa.RecordBatch.from_arrays([pa.array([1, 2, 3], t), pa.array([1, None, 2], t)], schema=schema)
    out, = oracle.project([root], [t], batch)
    assert out.to_pylist() == [1, None, 1]


def test_lineitem_generator_ranges(oracle):
    n = 100000
    ship, _ = oracle.generate_lineitem(0, 42, 0, n)
    disc, _ = oracle.generate_lineitem(1, 42, 0, n)
    qty, vld = oracle.generate_lineitem(2, 42, 0, n, null_permille=10)
    assert ship.min() >= 8035 and ship.max() <= 10561
    assert set(np.round(disc * 100).astype(int)) == set(range(11))
    assert qty.min() == 1.0 and qty.max() == 50.0
    nulls = n - int(np.unpackbits(vld[: (n + 7) // 8], bitorder="little")[:n].sum())
    assert 500 < nulls < 1500
    # row-range independence: generating a sub-range gives the same values
    part, _ = oracle.generate_lineitem(0, 42, 1000, 500)
    assert np.array_equal(part, ship[1000:1500])
    # Q6 selectivity of the synthetic data is ~1.8%
    sel = ((ship >= 8766) & (ship < 9131) & (disc >= 0.05) & (disc <= 0.07) & (qty < 24)).mean()
    assert 0.014 < sel < 0.023
