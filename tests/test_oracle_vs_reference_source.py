"""CPU, build container only: the numpy restatement (oracle/tpls_oracle.py) against the reference's UNMODIFIED
source executed here (/root/reference on top of the restated tensorly leaves, as oracle/make_golden.py does), on a
spread of seeded random configurations beyond the committed golden vectors -- orders 2 to 7, several responses, NaNs,
coupled tensors, fp32 storage.  Skipped where /root/reference does not exist (the GPU box)."""

import contextlib
import io
import os
import re
import sys
import warnings

import numpy as np
import pytest

from _util import aligned_errors

REFERENCE = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

pytestmark = pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="needs /root/reference (build container)")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle", "tensorly_standin"))
    sys.path.insert(0, REFERENCE)
    from cmtf_pls.tpls import tPLS
    from cmtf_pls.cmtf import ctPLS
    return tPLS, ctPLS


def _fit_reference(est, X, Y, max_iter=100):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        est.fit(X, Y, verbose=1, max_iter=max_iter)
    trips = np.full(est.n_components, max_iter, dtype=np.int64)
    for m in re.finditer(r"Comp (\d+): converged after (\d+) iterations", buf.getvalue()):
        trips[int(m.group(1))] = int(m.group(2)) + 1
    return trips


def _state(est, coupled, L):
    if coupled:
        return dict(T=est.factor_T, W=[est.Xs_factors[l][1:] for l in range(L)], U=est.Y_factors[0], Q=est.Y_factors[1],
                    coef=est.coef_, R2X=est.R2Xs, R2Y=est.R2Y)
    return dict(T=est.X_factors[0], W=[est.X_factors[1:]], U=est.Y_factors[0], Q=est.Y_factors[1], coef=est.coef_,
                R2X=[est.R2X], R2Y=est.R2Y)


SINGLE = [
    # shape, M, latent, R, error, nan fraction, dtype
    ((35, 9), 1, 3, 2, 0.2, 0.0, np.float64),
    ((35, 9), 3, 3, 3, 0.2, 0.1, np.float64),
    ((40, 7, 5), 2, 3, 3, 0.3, 0.0, np.float64),
    ((40, 7, 5), 4, 4, 3, 0.3, 0.15, np.float64),
    ((45, 6, 5, 4), 3, 4, 3, 0.4, 0.0, np.float64),
    ((45, 6, 5, 4), 3, 4, 2, 0.4, 0.1, np.float32),
    ((30, 4, 3, 3, 2), 2, 3, 2, 0.3, 0.0, np.float64),
    ((40, 4, 3, 3, 2, 2), 3, 3, 2, 0.5, 0.0, np.float64),       # 5-way covariance tensor
    ((36, 3, 3, 2, 2, 2, 2), 2, 3, 2, 0.5, 0.0, np.float64),    # 6-way covariance tensor
    ((30, 12, 10), 12, 5, 3, 0.3, 0.0, np.float64),             # more than 8 responses
]


@pytest.mark.parametrize("idx", range(len(SINGLE)))
def test_single_tensor_fit_equals_reference_source(ref, idx):
    from oracle import tpls_oracle as orc
    tPLS, _ = ref
    shape, M, L, R, err, nan, dtype = SINGLE[idx]
    X, Y, _ = orc.synthetic(shape, M, L, error=err, seed=1000 + idx)
    if nan:
        X[np.random.default_rng(idx).random(X.shape) < nan] = np.nan
    X = X.astype(dtype)
    est = tPLS(R)
    trips = _fit_reference(est, X.copy(), Y.copy())
    st = orc.fit([X.copy()], Y.copy(), R, r2_mode="reference")
    assert st["trips"].tolist() == trips.tolist()
    for k, e in aligned_errors(st, _state(est, False, 1)).items():
        assert e < 1e-9, (shape, k, e)


COUPLED = [
    ([(30, 6, 5), (30, 8)], 3, 3, 0.0),
    ([(32, 5, 4, 3), (32, 6, 5), (32, 7)], 4, 2, 0.1),
    ([(28, 4, 3, 2, 2), (28, 9)], 2, 2, 0.0),
]


@pytest.mark.parametrize("idx", range(len(COUPLED)))
def test_coupled_fit_equals_reference_source(ref, idx):
    from oracle import tpls_oracle as orc
    _, ctPLS = ref
    dims, M, R, nan = COUPLED[idx]
    rng = np.random.default_rng(50 + idx)
    n = dims[0][0]
    T = rng.normal(size=(n, 4))
    Xs = []
    for d in dims:
        facs = [T] + [rng.normal(size=(k, 4)) for k in d[1:]]
        X = orc.rank_r_tensor(facs) + rng.normal(0, 0.4, size=d)
        Xs.append(X)
    if nan:
        Xs[0][rng.random(Xs[0].shape) < nan] = np.nan
    Y = T @ rng.normal(size=(M, 4)).T + rng.normal(0, 0.3, size=(n, M))
    est = ctPLS(R)
    trips = _fit_reference(est, [x.copy() for x in Xs], Y.copy())
    st = orc.fit([x.copy() for x in Xs], Y.copy(), R, r2_mode="reference")
    assert st["trips"].tolist() == trips.tolist()
    for k, e in aligned_errors(st, _state(est, True, len(Xs))).items():
        assert e < 1e-9, (dims, k, e)
