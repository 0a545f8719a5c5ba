"""Known-answer vectors of the reference's descendant (tests/golden/gandiva_vectors.json,
transcribed from site-packages/pyarrow/tests/test_gandiva.py) replayed against
 (a) the CPU oracle                       -- pins the oracle      (CPU, always runs)
 (b) the CUDA path through the C-ABI      -- pins the product     (-m gpu)
"""
import json
import os

import numpy as np
import pyarrow as pa
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "gandiva_vectors.json")))
VECTORS = GOLD["vectors"]


def _type(name):
    return pa.type_for_alias(name)


def build_node(b, tree, schema):
    head = tree[0]
    if head == "field":
        if len(tree) == 3:
            return b.make_field(pa.field(tree[1], _type(tree[2])))
        return b.make_field(schema.field(tree[1]))
    if head == "lit":
        return b.make_literal(tree[1], _type(tree[2]))
    if head == "fn":
        return b.make_function(tree[1], [build_node(b, c, schema) for c in tree[3:]], _type(tree[2]))
    if head == "if":
        c, t, e = (build_node(b, x, schema) for x in tree[2:5])
        return b.make_if(c, t, e, _type(tree[1]))
    if head == "and":
        return b.make_and([build_node(b, c, schema) for c in tree[1:]])
    if head == "or":
        return b.make_or([build_node(b, c, schema) for c in tree[1:]])
    if head == "in":
        return b.make_in_expression(build_node(b, tree[2], schema), tree[3], _type(tree[1]))
    raise ValueError(head)


def make_batch(vec):
    schema = pa.schema([(n, _type(t)) for n, t in vec["schema"]])
    return pa.RecordBatch.from_arrays(
        [pa.array(vec["data"][f.name], type=f.type) for f in schema], schema=schema)


@pytest.mark.parametrize("vec", VECTORS, ids=[v["name"] for v in VECTORS])
def test_oracle_matches_golden(vec, gandiva, oracle):
    b = gandiva.TreeExprBuilder()
    batch = make_batch(vec)
    root = build_node(b, vec["expr"], batch.schema)
    if vec["kind"] == "project":
        rt = _type(vec["result_type"])
        out, = oracle.project([root], [rt], batch)
        assert out.to_pylist() == vec["expected"]
    elif vec["kind"] == "filter":
        idx = oracle.filter_indices(root, batch)
        assert idx.tolist() == vec["expected"]
    else:
        froot = build_node(b, vec["filter_expr"], batch.schema)
        idx = oracle.filter_indices(froot, batch)
        rt = _type(vec["result_type"])
        out, = oracle.project([root], [rt], batch, selection=idx)
        assert out.to_pylist() == vec["expected"]


@pytest.mark.gpu
@pytest.mark.parametrize("vec", VECTORS, ids=[v["name"] for v in VECTORS])
def test_gpu_matches_golden(vec, gandiva):
    b = gandiva.TreeExprBuilder()
    batch = make_batch(vec)
    root = build_node(b, vec["expr"], batch.schema)
    if vec["kind"] == "project":
        field = pa.field("res", _type(vec["result_type"]))
        p = gandiva.make_projector(batch.schema, [b.make_expression(root, field)],
                                   pa.default_memory_pool())
        r, = p.evaluate(batch)
        assert r.equals(pa.array(vec["expected"], type=field.type))
    elif vec["kind"] == "filter":
        f = gandiva.make_filter(batch.schema, b.make_condition(root))
        sel = f.evaluate(batch, pa.default_memory_pool())
        assert sel.to_array().equals(pa.array(vec["expected"], type=_type(vec["expected_type"])))
    else:
        froot = build_node(b, vec["filter_expr"], batch.schema)
        f = gandiva.make_filter(batch.schema, b.make_condition(froot))
        field = pa.field("res", _type(vec["result_type"]))
        p = gandiva.make_projector(batch.schema, [b.make_expression(root, field)],
                                   pa.default_memory_pool(), vec["selection_mode"])
        sel = f.evaluate(batch, pa.default_memory_pool())
        r, = p.evaluate(batch, sel)
        assert r.equals(pa.array(vec["expected"], type=field.type))


def test_to_string_formats(gandiva):
    """ToString pins (test_gandiva.py:376-393)."""
    b = gandiva.TreeExprBuilder()
    for case in GOLD["to_string"]["cases"]:
        s = str(build_node(b, case["node"], None))
        if "equals" in case:
            assert s == case["equals"]
        else:
            assert s.startswith(case["startswith"])
