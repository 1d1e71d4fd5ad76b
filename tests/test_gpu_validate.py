"""GPU: the cross-validation sweep -- folds as 0/1 row weights,
one R-component fit per fold -- against the oracle's refits on sliced copies."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,coupled,miss", [((90, 12, 8), False, False), ((80, 8, 6, 4), False, False),
                                                ((90, 12, 8), True, False), ((90, 12, 8), False, True)])
def test_q2y_sweep_matches_oracle_refits(shape, coupled, miss):
    from oracle import tpls_oracle as orc
    from cmtf_pls_b200 import q2y_sweep
    from cmtf_pls_b200.validate import _folds
    X, Y, _ = orc.synthetic(shape, 3, 4, error=0.4, seed=9, extra_dims=[(shape[0], 10)] if coupled else ())
    if miss:
        rng = np.random.default_rng(2)
        X[rng.random(X.shape) < 0.1] = np.nan
    folds = _folds(shape[0], 5, 3)
    q_ref, cv_ref = orc.q2y_kfold(X, Y, 4, folds)
    for alg in ("stream", "covariance"):
        q, cv = q2y_sweep(X, Y, 4, n_splits=5, seed=3, return_scores=True, algorithm=alg)
        assert np.max(np.abs(q - q_ref)) < 1e-8, alg
        # held-out rows' scores are transform() of the held-out data under each fold's model
        s = np.sign(np.sum(cv * cv_ref, axis=0))
        assert np.max(np.abs(cv * s - cv_ref)) / np.max(np.abs(cv_ref)) < 1e-8, alg
        assert q[-1] > 0.5 and np.all(np.diff(q[:3]) > -1e-3)


def test_get_q2y_leave_one_out_like_reference():
    """validate.py:24-37 restated: LOO refits, uncentred denominator."""
    from oracle import tpls_oracle as orc
    from cmtf_pls_b200 import tPLS, get_q2y
    X, Y, _ = orc.synthetic((14, 6, 5), 2, 2, error=0.2, seed=4)
    pls = tPLS(2)
    pls.fit(X, Y)
    q_ref, _ = orc.q2y_kfold(X, Y, 2, [np.array([i]) for i in range(14)])
    assert abs(get_q2y(pls) - q_ref[-1]) < 1e-8


def test_weights_of_one_change_nothing():
    from oracle import tpls_oracle as orc
    from cmtf_pls_b200 import _core
    X, Y, _ = orc.synthetic((60, 10, 6), 3, 4, error=0.4, seed=1)
    a = _core.run_fit([X], Y, 3, 1e-8, 100)
    b = _core.run_fit([X], Y, 3, 1e-8, 100, row_weights=np.ones(60))
    for k in ("T", "U", "Q", "coef", "R2Y"):
        assert np.max(np.abs(a[k] - b[k])) < 1e-12, k
    assert np.max(np.abs(a["R2X"][0] - b["R2X"][0])) < 1e-12


def test_cv_fold_with_all_nans_in_held_out_rows_stays_finite():
    """the missing-data flag must come from an UNWEIGHTED NaN census -- a fold whose NaNs all sit in
    held-out rows still needs the masked kernels (the dense ones would turn 0 * NaN into NaN everywhere)."""
    from oracle import tpls_oracle as orc
    from cmtf_pls_b200 import q2y_sweep
    from cmtf_pls_b200.validate import _folds
    X, Y, _ = orc.synthetic((60, 10, 6), 3, 4, error=0.3, seed=31)
    folds = _folds(60, 5, 1)
    X[folds[2][0], 3, 2] = np.nan          # the only NaNs sit in ONE sample, held out by fold 2
    X[folds[2][0], 7, 1] = np.nan
    q_ref, cv_ref = orc.q2y_kfold(X, Y, 3, folds)
    for alg in ("stream", "covariance"):
        q, cv = q2y_sweep(X, Y, 3, folds=folds, return_scores=True, algorithm=alg)
        assert np.all(np.isfinite(q)) and np.all(np.isfinite(cv))
        assert np.max(np.abs(q - q_ref)) < 1e-8, alg


def test_covariance_loop_with_three_coupled_tensors_matches_oracle():
    """The covariance-mode inner loop runs one CTA per coupled tensor in a thread-block cluster (the per-tensor parts
    of Y't meet through distributed shared memory): three tensors of different order, against the oracle."""
    from oracle import tpls_oracle as orc
    from cmtf_pls_b200 import ctPLS
    Xs, Y, _ = orc.synthetic((120, 16, 8), 4, 5, error=0.4, seed=17, extra_dims=[(120, 6, 5, 4), (120, 20)])
    ref = orc.fit([x.copy() for x in Xs], Y.copy(), 4, r2_mode="residual")
    for alg in ("covariance", "stream"):
        est = ctPLS(4, algorithm=alg)
        est.fit([x.copy() for x in Xs], Y.copy())
        assert bool(est.stats_["covariance_mode"]) == (alg == "covariance")
        assert est.n_iter_.tolist() == ref["trips"].tolist(), alg
        s = np.sign(np.sum(est.factor_T * ref["T"], axis=0))
        assert np.max(np.abs(est.factor_T * s - ref["T"])) / np.max(np.abs(ref["T"])) < 1e-8, alg
        assert np.max(np.abs(est.R2Y - ref["R2Y"])) < 1e-8, alg
        for l in range(3):
            assert np.max(np.abs(est.R2Xs[l] - ref["R2X"][l])) < 1e-8, (alg, l)
