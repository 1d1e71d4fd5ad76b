"""CPU: host-side plumbing of the estimators that needs no GPU -- the deferred
profile read-out, the result-buffer policy, the fold partition of the
cross-validation sweep, and the y-score recurrence of transform."""

import numpy as np


def test_lazy_profile_reads_the_engine_once():
    from cmtf_pls_b200._core import LazyProfile

    class FakeEngine:
        calls = 0

        def profile(self):
            FakeEngine.calls += 1
            return {"contract": dict(ms=1.0, launches=2, bytes=3.0)}

    p = LazyProfile(FakeEngine())
    assert FakeEngine.calls == 0                 # nothing is queried inside fit
    assert p.get()["contract"]["launches"] == 2
    assert p.get() is p.get() and FakeEngine.calls == 1


def test_result_buffers_first_fit_is_pageable():
    """The first two large fetches of an engine (T and U of the first fit) never pin memory; small results never do."""
    from cmtf_pls_b200._engine import Engine
    eng = Engine.__new__(Engine)                 # no device needed for the buffer policy
    a = eng._host_out(600_000, 1)                # 4.8 MB: counted
    b = eng._host_out(600_000, 1)
    c = eng._host_out(10, 3)                     # small: not counted
    assert a.shape == (600_000, 1) and a.dtype == np.float64 and c.shape == (10, 3)
    assert eng._big_fetches == 2 and a.flags["C_CONTIGUOUS"] and b.flags["WRITEABLE"]
    d = eng._host_out(600_000, 1)                # from here on pinned when a CUDA torch is there, else pageable
    assert d.shape == (600_000, 1) and eng._big_fetches == 3


def test_folds_partition_the_samples():
    from cmtf_pls_b200.validate import _folds
    folds = _folds(103, 5, seed=3)
    assert len(folds) == 5 and sorted(np.concatenate(folds).tolist()) == list(range(103))
    assert all(np.all(np.diff(f) > 0) for f in folds)
    assert [f.tolist() for f in _folds(4, None, 0)] == [[0], [1], [2], [3]]        # leave-one-out, in order
    again = _folds(103, 5, seed=3)
    assert all(np.array_equal(a, b) for a, b in zip(folds, again))


def test_y_scores_follow_the_reference_recurrence():
    """tpls.py:179-184: U[:, a] = Y q_a, then Y -= T coef[:, [a]] q_a^T."""
    from cmtf_pls_b200._core import y_scores
    rng = np.random.default_rng(0)
    n, m, R = 12, 3, 2
    Y, T = rng.normal(size=(n, m)), rng.normal(size=(n, R))
    coef, Q, mu = np.triu(rng.normal(size=(R, R))), rng.normal(size=(m, R)), rng.normal(size=m)
    got = y_scores(Y, mu, (n, m), T, coef, Q)
    Yc = Y - mu
    want = np.zeros((n, R))
    for a in range(R):
        want[:, a] = Yc @ Q[:, a]
        Yc = Yc - T @ coef[:, [a]] @ Q[:, [a]].T
    assert np.allclose(got, want, atol=1e-14)
    import pytest
    with pytest.raises(ValueError, match="Training Y has shape"):
        y_scores(rng.normal(size=(n, m + 1)), mu, (n, m), T, coef, Q)
