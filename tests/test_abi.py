"""CPU: the C-ABI library builds in-tree, loads, and exports every symbol that
include/tpls_b200.h declares (no compute calls without a GPU)."""

import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "tpls_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tpls_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from cmtf_pls_b200 import _engine
    if not os.path.exists(_engine.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = _engine.load_library()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_engine.SIGNATURES) == names
    assert lib.tpls_version() >= 100


def test_no_cpu_fallback():
    """Without a CUDA device the estimators must fail loudly, not compute."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cmtf_pls_b200 import tPLS
    from cmtf_pls_b200._engine import TplsError
    with pytest.raises(TplsError, match="no CUDA device|no CPU"):
        tPLS(2).fit(np.zeros((5, 4, 3)), np.zeros((5, 2)))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "cmtf_pls_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "tensorly" not in text or f.endswith((".cu", ".cuh")), f


def test_argument_checks_match_reference():
    """AssertionError on sample-count mismatch / 3-D Y / non-list Xs (tpls.py:46-47, cmtf.py:46-51)."""
    import numpy as np
    from cmtf_pls_b200 import tPLS, ctPLS
    with pytest.raises(AssertionError):
        tPLS(2).fit(np.zeros((5, 4, 3)), np.zeros((6, 2)))
    with pytest.raises(AssertionError):
        tPLS(2).fit(np.zeros((5, 4, 3)), np.zeros((5, 2, 2)))
    with pytest.raises(AssertionError):
        ctPLS(2).fit(np.zeros((5, 4, 3)), np.zeros((5, 2)))
    with pytest.raises(AssertionError):
        ctPLS(2).fit([np.zeros((5,))], np.zeros((5, 2)))
    p = tPLS(2)
    assert len(p) == 3
    with pytest.raises(IndexError):
        p[3]
