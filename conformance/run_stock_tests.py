"""Runs the UNMODIFIED binding tests of the reference's descendant
(site-packages/pyarrow/tests/test_gandiva.py) against this engine.

How: the stock Cython module `pyarrow/gandiva.pyx` is compiled, unmodified, against
include/gandiva/*.h and linked to gandiva_b200/libgandiva.so (gandiva_b200/build.py:
build_stock_binding), then registered as `pyarrow.gandiva` for this process only.
Expected divergences: none in assertions about results; the two
`llvm_ir.find("@expr_") != -1` assertions (test_tree_exp_builder, test_filter) FAIL: DumpIR
returns the generated CUDA source + PTX (kernels gdv_project_expr_<hash> / gdv_filter_expr_<hash>),
not LLVM IR, and the output is not shaped to contain "@expr_".  tests/test_stock_conformance.py
asserts exactly that outcome: 9 passed, 1 skipped upstream, those 2 failed.

  python conformance/run_stock_tests.py [pytest args]        (needs a GPU for the evaluate tests)
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def load_stock_module():
    import pyarrow  # noqa: F401  (libarrow must be loaded first)
    stock_dir = os.path.join(HERE, "stock")
    cands = [f for f in os.listdir(stock_dir) if f.startswith("gandiva.") and f.endswith(".so")]
    if not cands:
        raise ImportError("conformance/stock/gandiva*.so missing: run python -m gandiva_b200.build")
    spec = importlib.util.spec_from_file_location("pyarrow.gandiva", os.path.join(stock_dir, cands[0]))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["pyarrow.gandiva"] = mod
    spec.loader.exec_module(mod)
    import pyarrow
    pyarrow.gandiva = mod
    return mod


def main(argv):
    import pyarrow
    import pytest
    load_stock_module()
    test_file = os.path.join(os.path.dirname(pyarrow.__file__), "tests", "test_gandiva.py")
    return pytest.main([test_file, "-q", "-p", "no:cacheprovider", "-W", "ignore::FutureWarning",
                        "-W", "ignore::DeprecationWarning"] + argv)


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
