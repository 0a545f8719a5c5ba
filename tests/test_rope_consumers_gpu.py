"""Consumers of a rope (csrc/gdv_rope_temps.h).  concat / repeat / lpad / rpad / reverse / replace results are ropes
that only projection, concat and if/else read directly; any other consumer — like, equal, substr, IN, a length over a
padded string ... — is served by materialising the rope into a temporary utf8 column first (an internal Projector)
and evaluating the caller's expression over the batch plus that column.  Bit-exact against the oracle, which has no
such distinction (it is a scalar interpreter over std::string)."""
import numpy as np
import pyarrow as pa
import pytest

import cases
import devmem
from helpers import assert_arrays_match

pytestmark = pytest.mark.gpu

S, B, I32, I64 = pa.string(), pa.bool_(), pa.int32(), pa.int64()
SCHEMA = pa.schema([("s", S), ("u", S), ("k", I32)])


def _batch(n, seed, null_prob=0.1, offset=0):
    rng = np.random.default_rng(seed)
    alphabet = list("abcxyzAB 01_") + ["é", "日"]
    mk = lambda: pa.array([None if rng.random() < null_prob else "".join(rng.choice(alphabet, size=int(rng.integers(0, 14))))
                           for _ in range(n + offset)], S)
    cols = [mk(), mk(), pa.array(rng.integers(0, 9, n + offset, dtype=np.int32), mask=rng.random(n + offset) < null_prob)]
    if offset:
        cols = [c.slice(offset) for c in cols]
    return pa.RecordBatch.from_arrays(cols, schema=SCHEMA)


def _roots(b):
    f = {x.name: b.make_field(x) for x in SCHEMA}
    fn = b.make_function
    lit = lambda v, t=S: b.make_literal(v, t)
    cat = fn("concat", [f["s"], lit(" "), f["u"]], S)
    catop = fn("concatOperator", [f["s"], f["u"]], S)
    return f, [
        (fn("like", [cat, lit("%ab%")], B), B),
        (fn("equal", [fn("reverse", [f["s"]], S), f["u"]], B), B),
        (fn("char_length", [fn("lpad", [f["s"], lit(10, I32), lit("xy")], S)], I32), I32),
        (fn("substr", [cat, lit(2, I64), lit(6, I64)], S), S),
        (fn("upper", [fn("reverse", [f["s"]], S)], S), S),
        (b.make_in_expression(catop, ["ab", "a日", "", "abab"], S), B),
        (fn("starts_with", [fn("replace", [f["s"], lit("a"), lit("bb")], S), lit("bb")], B), B),
        (fn("locate", [lit("b"), fn("rpad", [f["u"], lit(6, I32), lit("ab")], S)], I32), I32),
        (fn("concat", [fn("substr", [cat, lit(1, I64), lit(3, I64)], S), lit("|"), fn("reverse", [f["u"]], S)], S), S),
        (b.make_if(fn("like", [catop, lit("a%")], B), fn("repeat", [f["s"], lit(2, I32)], S), f["u"], S), S),
    ]


@pytest.mark.parametrize("n,offset", [(1, 0), (777, 0), (5000, 3)])
def test_projector_over_materialised_ropes(n, offset, gandiva, oracle):
    b = gandiva.TreeExprBuilder()
    _, roots = _roots(b)
    batch = _batch(n, seed=n, offset=offset)
    for k, (root, t) in enumerate(roots):
        p = gandiva.make_projector(SCHEMA, [b.make_expression(root, pa.field("o", t))], None)
        got = p.evaluate(batch)[0]
        want = oracle.project([root], [t], batch)[0]
        assert_arrays_match(got, want, "rope consumer %d: %s" % (k, root))
    # several outputs at once share the temps (one materialisation per distinct rope)
    p = gandiva.make_projector(SCHEMA, [b.make_expression(r, pa.field("o%d" % k, t)) for k, (r, t) in enumerate(roots)], None)
    got = p.evaluate(batch)
    want = oracle.project([r for r, _ in roots], [t for _, t in roots], batch)
    for k, (gv, wv) in enumerate(zip(got, want)):
        assert_arrays_match(gv, wv, "all outputs, %d" % k)


def test_filter_over_materialised_ropes(gandiva, oracle):
    b = gandiva.TreeExprBuilder()
    f, roots = _roots(b)
    batch = _batch(6000, seed=9)
    conds = [r for r, t in roots if t == B]
    conds.append(b.make_and([conds[0], b.make_function("less_than", [f["k"], b.make_literal(5, I32)], B)]))
    conds.append(b.make_or([conds[1], conds[2]]))
    for k, cond in enumerate(conds):
        flt = gandiva.make_filter(SCHEMA, b.make_condition(cond))
        got = flt.evaluate(batch).to_array().to_numpy().astype(np.uint64)
        want = oracle.filter_indices(cond, batch)
        assert np.array_equal(got, want), "filter %d: %s" % (k, cond)
        assert 0 <= len(want) <= 6000


def test_projector_with_selection_vector_over_ropes(gandiva, oracle):
    b = gandiva.TreeExprBuilder()
    _, roots = _roots(b)
    batch = _batch(4000, seed=5)
    sel_idx = np.sort(np.random.default_rng(1).choice(4000, size=900, replace=False)).astype(np.uint32)
    for root, t in (roots[0], roots[3]):
        p = gandiva.make_projector(SCHEMA, [b.make_expression(root, pa.field("o", t))], None, "UINT32")
        sv = gandiva.SelectionVector(sel_idx, len(sel_idx), gandiva._SEL_MODE["UINT32"])
        got = p.evaluate(batch, sv)[0]
        want = oracle.project([root], [t], batch, selection=sel_idx.astype(np.int64))[0]
        assert len(got) == len(sel_idx)
        assert_arrays_match(got, want, "selection vector over a rope consumer")


def test_the_two_stage_plan_is_what_ran(gandiva):
    """DumpIR shows the temporary column the consumer reads, and an expression the fuser takes directly is untouched."""
    b = gandiva.TreeExprBuilder()
    f = {x.name: b.make_field(x) for x in SCHEMA}
    fn = b.make_function
    cat = fn("concat", [f["s"], f["u"]], S)
    cfg = gandiva.Configuration(dump_ir=True)
    p = gandiva.make_projector(SCHEMA, [b.make_expression(fn("like", [cat, b.make_literal("%ab%", S)], B), pa.field("o", B))],
                               None, configuration=cfg)
    assert "__gdv_rope_0" in p.llvm_ir
    q = gandiva.make_projector(SCHEMA, [b.make_expression(cat, pa.field("o", S))], None, configuration=cfg)
    assert "__gdv_rope_" not in q.llvm_ir


def test_device_resident_batch_over_ropes(gandiva, oracle):
    """Filter over a batch that lives in device memory: the temporary column is allocated from the engine's pool."""
    n = 3000
    batch = _batch(n, seed=31, null_prob=0.0)
    b = gandiva.TreeExprBuilder()
    f = {x.name: b.make_field(x) for x in SCHEMA}
    fn = b.make_function
    cond = fn("like", [fn("concat", [f["s"], b.make_literal("-", S), f["u"]], S), b.make_literal("%a-%", S)], B)
    flt = gandiva.make_filter(SCHEMA, b.make_condition(cond))
    bufs = []

    def dev(arr_buf, dtype):
        a = np.frombuffer(arr_buf, dtype=dtype)
        d = devmem.DevBuf(len(a), dtype)
        if devmem.EMU:
            d.a[...] = a
        else:
            import torch
            d.a.copy_(torch.from_numpy(a.copy()))
        bufs.append(d)
        return d.ptr
    cols = []
    for name in ("s", "u"):
        arr = batch.column(name)
        offs, data = arr.buffers()[1], arr.buffers()[2]
        cols.append((0, dev(offs, np.int32), dev(data, np.uint8) if data is not None and data.size else 0, 0))
    karr = batch.column("k")
    cols.append((0, dev(karr.buffers()[1], np.int32), 0, 0))
    out = devmem.DevBuf(n, np.int64)
    cnt = devmem.DevBuf(1, np.int64, fill=0)
    got_n = flt.evaluate_device(n, cols, out.ptr, n, "UINT64", devmem.stream(), cnt.ptr, sync=True)
    want = oracle.filter_indices(cond, batch)
    assert got_n == len(want)
    assert np.array_equal(out.numpy()[:got_n].astype(np.uint64), want)


def test_nested_rope_consumers(gandiva, oracle):
    """A consumer inside the arguments of a rope that is itself consumed: one level of temporaries per nesting level."""
    b = gandiva.TreeExprBuilder()
    f = {x.name: b.make_field(x) for x in SCHEMA}
    fn = b.make_function
    inner = fn("substr", [fn("reverse", [f["s"]], S), b.make_literal(1, I64), b.make_literal(3, I64)], S)
    deeper = fn("upper", [fn("lpad", [fn("substr", [fn("concat", [inner, f["u"]], S), b.make_literal(2, I64)], S),
                                      b.make_literal(8, I32), b.make_literal("*", S)], S)], S)
    roots = [(fn("like", [fn("concat", [inner, f["u"]], S), b.make_literal("a%", S)], B), B), (deeper, S)]
    batch = _batch(2500, seed=77)
    for root, t in roots:
        p = gandiva.make_projector(SCHEMA, [b.make_expression(root, pa.field("o", t))], None)
        assert_arrays_match(p.evaluate(batch)[0], oracle.project([root], [t], batch)[0], str(root))
    flt = gandiva.make_filter(SCHEMA, b.make_condition(roots[0][0]))
    assert np.array_equal(flt.evaluate(batch).to_array().to_numpy().astype(np.uint64), oracle.filter_indices(roots[0][0], batch))


def test_concat_wider_than_one_rope(gandiva, oracle):
    """A rope holds 8 pieces.  A concat of more arguments (or of ropes that add up to more) reads its widest rope
    arguments through temporaries, and a concat of more than 8 scalar arguments is folded 8 at a time — the same
    strings as the oracle's plain left-to-right concatenation, NULL rules of concat / concatOperator included."""
    b = gandiva.TreeExprBuilder()
    f = {x.name: b.make_field(x) for x in SCHEMA}
    fn = b.make_function
    lit = lambda v, t=S: b.make_literal(v, t)
    ten = [f["s"], lit("-"), f["u"], lit("+"), f["s"], f["u"], lit("日"), f["s"], lit(""), f["u"]]   # the registry's widest concat
    nested = fn("concat", [fn("concatOperator", [f["s"], f["u"], f["s"], f["u"], f["s"]], S),
                           fn("lpad", [f["u"], lit(6, I32), lit("ab")], S), fn("reverse", [f["s"]], S),
                           fn("concat", [f["u"], lit("/"), f["s"], lit("/"), f["u"]], S), f["s"]], S)
    roots = [(fn("concat", ten, S), S), (fn("concatOperator", ten, S), S), (nested, S),
             (fn("char_length", [fn("concat", [fn("concat", ten, S), lit("|"), fn("concatOperator", ten, S)], S)], I32), I32),
             (fn("like", [fn("concatOperator", ten, S), lit("%a-%")], B), B)]
    batch = _batch(3000, seed=41, null_prob=0.08)
    for root, t in roots:
        p = gandiva.make_projector(SCHEMA, [b.make_expression(root, pa.field("o", t))], None)
        assert_arrays_match(p.evaluate(batch)[0], oracle.project([root], [t], batch)[0], str(root)[:80])
