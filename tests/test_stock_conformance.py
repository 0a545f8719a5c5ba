"""The reference's own (descendant's) binding tests, UNMODIFIED, against this engine: the stock
Cython module pyarrow/gandiva.pyx is compiled against include/gandiva/*.h and run through
site-packages/pyarrow/tests/test_gandiva.py (conformance/run_stock_tests.py)."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNNER = os.path.join(ROOT, "conformance", "run_stock_tests.py")


def _stock_built():
    d = os.path.join(ROOT, "conformance", "stock")
    return os.path.isdir(d) and any(f.endswith(".so") for f in os.listdir(d))


def _run(extra):
    res = subprocess.run([sys.executable, RUNNER] + extra, capture_output=True, text=True, timeout=600)
    return res.returncode, res.stdout + res.stderr


@pytest.mark.skipif(not _stock_built(), reason="stock binding not built (needs Cython + pyarrow headers)")
def test_stock_host_side_tests_pass_without_gpu():
    """Tests that never call Evaluate: literals, ToString formats, None rejection, registry."""
    rc, out = _run(["-k", "literals or to_string or rejects_none or registered"])
    assert rc == 0, out[-3000:]
    assert re.search(r"4 passed", out), out[-2000:]


@pytest.mark.gpu
@pytest.mark.skipif(not _stock_built(), reason="stock binding not built")
def test_stock_binding_tests_pass_on_gpu():
    """test_gandiva.py on the CUDA path: every test that checks RESULTS passes (9 + 1 skipped
    upstream).  The two known divergences are the assertions that DumpIR
    contains LLVM's "@expr_" symbol: here DumpIR is CUDA source + PTX (kernel
    gdv_project_expr_N / gdv_filter_expr_N); output is not shaped to satisfy them."""
    rc, out = _run([])
    assert re.search(r"2 failed, 9 passed, 1 skipped", out), out[-4000:]
    failed = re.findall(r"^FAILED .*::(\w+)", out, flags=re.M)
    assert sorted(failed) == ["test_filter", "test_tree_exp_builder"], failed
    # both failures are the llvm_ir "@expr_" assertion and nothing else
    assert len(re.findall(r'llvm_ir\.find\("@expr_"\) != -1', out)) >= 2
    assert "Error" not in "".join(l for l in out.splitlines() if l.startswith("E  ") and "AssertionError" not in l and "assert -1" not in l and "where" not in l)
