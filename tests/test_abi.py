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


def test_stats_struct_of_the_binding_matches_the_header():
    """`tpls_stats` (include/tpls_b200.h) field by field against the ctypes mirror: same names, order and types -- the
    struct is filled by the library and read through ctypes, so a field added on one side only shifts everything after it."""
    import ctypes as C
    import re
    from cmtf_pls_b200 import _engine
    src = open(os.path.join(ROOT, "include", "tpls_b200.h")).read()
    body = re.search(r"typedef struct tpls_stats \{(.*?)\} tpls_stats;", src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"(double|int64_t|int)\s+(\w+)\s*;", body)
    ctype = {"double": C.c_double, "int64_t": C.c_int64, "int": C.c_int}
    assert [(n, ctype[t]) for t, n in fields] == list(_engine.Stats._fields_)
    assert "working sets" in src and "256" in src      # (the header states the resident loop's default size limit)
