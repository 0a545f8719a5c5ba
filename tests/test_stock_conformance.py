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
    """All of test_gandiva.py (11 tests + 1 skipped upstream) passes on the CUDA path."""
    rc, out = _run([])
    assert rc == 0, out[-4000:]
    assert re.search(r"11 passed, 1 skipped", out), out[-2000:]
