"""Writes tests/golden/gandiva_vectors.json.

The reference mount has no source or tests and no Gandiva build exists in this
image, so the only known-answer vectors for this path are the ones the reference's
maintained descendant ships as text: site-packages/pyarrow/tests/test_gandiva.py (Apache
Arrow 24.0.0).  They are transcribed here by hand, each with the lines it comes from; the
`expected` values are the ones asserted there, not outputs of any code in this repo.

Run: python tests/golden/make_golden.py
"""
import json
import os

SRC = "site-packages/pyarrow/tests/test_gandiva.py"

VECTORS = [
    {"name": "if_greater_int32", "source": SRC + ":25-63", "kind": "project",
     "schema": [["a", "int32"], ["b", "int32"]],
     "data": {"a": [10, 12, -20, 5], "b": [5, 15, 15, 17]},
     "expr": ["if", "int32", ["fn", "greater_than", "bool", ["field", "a"], ["field", "b"]],
              ["field", "a"], ["field", "b"]],
     "result_type": "int32", "expected": [10, 15, 15, 17]},
    {"name": "add_float64", "source": SRC + ":66-90", "kind": "project",
     "schema": [["a", "double"], ["b", "double"]],
     "data": {"a": [1.0, 2.0], "b": [3.0, 4.0]},
     "expr": ["fn", "add", "double", ["field", "a"], ["field", "b"]],
     "result_type": "double", "expected": [4.0, 6.0]},
    {"name": "filter_less_than_1000", "source": SRC + ":93-114", "kind": "filter",
     "schema": [["a", "double"]],
     "data": {"a": [1.0 * i for i in range(10000)]},
     "expr": ["fn", "less_than", "bool", ["field", "a"], ["lit", 1000.0, "double"]],
     "expected": list(range(1000)), "expected_type": "uint32"},
    {"name": "in_string", "source": SRC + ":117-130", "kind": "filter",
     "schema": [["a", "string"]],
     "data": {"a": ["ga", "an", "nd", "di", "iv", "va"]},
     "expr": ["in", "string", ["field", "a"], ["an", "nd"]],
     "expected": [1, 2], "expected_type": "uint32"},
    {"name": "in_int32", "source": SRC + ":132-140", "kind": "filter",
     "schema": [["a", "int32"]],
     "data": {"a": [3, 1, 4, 1, 5, 9, 2, 6, 5, 4]},
     "expr": ["in", "int32", ["field", "a"], [1, 5]],
     "expected": [1, 3, 4, 8], "expected_type": "uint32"},
    {"name": "in_int64", "source": SRC + ":142-151", "kind": "filter",
     "schema": [["a", "int64"]],
     "data": {"a": [3, 1, 4, 1, 5, 9, 2, 6, 5, 4]},
     "expr": ["in", "int64", ["field", "a"], [1, 5]],
     "expected": [1, 3, 4, 8], "expected_type": "uint32"},
    {"name": "boolean_and_or", "source": SRC + ":228-252", "kind": "filter",
     "schema": [["a", "double"], ["b", "double"]],
     "data": {"a": [1., 31., 46., 3., 57., 44., 22.], "b": [5., 45., 36., 73., 83., 23., 76.]},
     "expr": ["or",
              ["and", ["fn", "less_than", "bool", ["field", "a"], ["lit", 50.0, "double"]],
               ["fn", "greater_than", "bool", ["field", "a"], ["field", "b"]]],
              ["fn", "less_than", "bool", ["field", "b"], ["lit", 11.0, "double"]]],
     "expected": [0, 2, 5], "expected_type": "uint32"},
    {"name": "like_spark", "source": SRC + ":295-316", "kind": "project",
     "schema": [["a", "string"]],
     "data": {"a": ["park", "sparkle", "bright spark and fire", "spark"]},
     "expr": ["fn", "like", "bool", ["field", "a"], ["lit", "%spark%", "string"]],
     "result_type": "bool", "expected": [False, True, True, True]},
    {"name": "filter_then_project_selection", "source": SRC + ":329-373", "kind": "filter_project",
     "schema": [["a", "int32"], ["b", "int32"], ["c", "int32"]],
     "data": {"a": [10, 12, -20, 5, 21, 29], "b": [5, 15, 15, 17, 12, 3],
              "c": [1, 25, 11, 30, -21, None]},
     "filter_expr": ["fn", "greater_than", "bool", ["field", "a"], ["field", "b"]],
     "expr": ["if", "int32", ["fn", "less_than", "bool", ["field", "b"], ["field", "c"]],
              ["field", "b"], ["field", "c"]],
     "result_type": "int32", "selection_mode": "UINT32", "expected": [1, -21, None]},
]

TO_STRING = {
    "source": SRC + ":376-393",
    "cases": [
        {"node": ["lit", 2.0, "double"], "startswith": "(const double) 2 raw("},
        {"node": ["lit", 2, "int64"], "equals": "(const int64) 2"},
        {"node": ["field", "x", "double"], "equals": "(double) x"},
        {"node": ["field", "y", "string"], "equals": "(string) y"},
        {"node": ["fn", "not", "bool", ["field", "z", "bool"]], "equals": "bool not((bool) z)"},
        {"node": ["and", ["fn", "not", "bool", ["field", "z", "bool"]], ["field", "y", "bool"]],
         "equals": "bool not((bool) z) && (bool) y"},
    ],
}

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gandiva_vectors.json")
    with open(out, "w") as f:
        json.dump({"vectors": VECTORS, "to_string": TO_STRING}, f, indent=1)
    print("wrote", out, len(VECTORS), "vectors")
