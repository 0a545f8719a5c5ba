"""oracle — CPU restatement of the reference path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this package; the product package `gandiva_b200` never does.

PARITY UNPINNED outside the eight known-answer vectors of
site-packages/pyarrow/tests/test_gandiva.py (see gdv_oracle.cc header and DESIGN.md):
/root/reference contains no source, so there is nothing to compile into oracle/_ref.
"""
from .gdv_oracle import (filter_indices, generate_lineitem, hardware_threads, project,  # noqa: F401
                         sexpr)
