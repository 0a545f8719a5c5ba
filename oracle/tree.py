"""A product-free stand-in for the TreeExprBuilder, for the CPU legs of bench.py.

TEST / BENCH INFRASTRUCTURE ONLY.  tests/cases.py builds expression trees through any object with the
`make_*` methods of `pyarrow.gandiva.TreeExprBuilder` (P/gandiva.pyx:283-589); the oracle serialises
the nodes' `kind / dtype / children / payload` (oracle/gdv_oracle.py::sexpr).  This builder makes
exactly those plain nodes, so `bench.py --impl reference` can hand the Q6 condition to the oracle
without loading libgandiva_b200.so."""
from __future__ import annotations

import decimal
from typing import Any, Sequence

import pyarrow as pa


class Node:
    def __init__(self, kind: str, dtype: pa.DataType, children: Sequence["Node"] = (), payload: Any = None):
        self.kind, self.dtype, self.children, self.payload = kind, dtype, list(children), payload

    def return_type(self) -> pa.DataType:
        return self.dtype


class TreeBuilder:
    def make_field(self, field: pa.Field) -> Node:
        return Node("field", field.type, payload=field.name)

    def make_literal(self, value: Any, dtype: pa.DataType) -> Node:
        if value is not None:
            if pa.types.is_string(dtype) or pa.types.is_binary(dtype):
                value = value.encode("utf-8") if isinstance(value, str) else bytes(value)
            elif pa.types.is_decimal128(dtype) and not isinstance(value, int):
                value = int(decimal.Decimal(value).scaleb(dtype.scale).to_integral_value())
        return Node("literal", dtype, payload=value)

    def make_function(self, name: str, children: Sequence[Node], return_type: pa.DataType) -> Node:
        return Node("function", return_type, children, name)

    def make_if(self, cond: Node, then: Node, otherwise: Node, return_type: pa.DataType) -> Node:
        return Node("if", return_type, [cond, then, otherwise])

    def make_and(self, children: Sequence[Node]) -> Node:
        return Node("and", pa.bool_(), children)

    def make_or(self, children: Sequence[Node]) -> Node:
        return Node("or", pa.bool_(), children)

    def make_in_expression(self, node: Node, values: Sequence[Any], dtype: pa.DataType) -> Node:
        return Node("in", pa.bool_(), [node], (dtype, list(values)))
