"""TEST INFRASTRUCTURE ONLY -- CPU (numpy) restatement of the tensor-PLS hot
path of meyer-lab/cmtf-pls.  It is the *checker* for the CUDA path and the
``cpu_baseline`` of bench.py; nothing in ``cmtf_pls_b200`` may import it.

What is restated, and where it lives in the reference:
  * ``tPLS.fit``  cmtf_pls/tpls.py:73-120   and ``ctPLS.fit``  cmtf.py:85-140
    -> :func:`fit` (one coupled engine; a single tensor is the L=1 case, which
    the reference's own ``ctPLS([X])`` reproduces to <= 5e-15)
  * ``preprocess``  tpls.py:44-71, cmtf.py:44-83          -> :func:`_centre`
  * ``miss_tensordot`` / ``miss_mmodedot``  missingvals.py:7-38
                                 -> :func:`contract` / :func:`project`
  * ``calcR2X`` util.py:7-15, ``factors_to_tensor`` util.py:18-20
                                 -> :func:`r2` / :func:`rank_r_tensor`
  * ``transform`` / ``predict``  tpls.py:122-186, cmtf.py:142-231
                                 -> :func:`transform` / :func:`predict`
  * ``import_synthetic`` draw order  synthetic.py:59-77 -> :func:`synthetic`
The rank-1 step (tpls.py:84-88) calls tensorly's ``parafac``; its restatement
is oracle/tensorly_standin/tensorly/decomposition/_cp.py.

PINNING: tests/golden/*.npz were produced by oracle/make_golden.py, which runs
the reference's UNMODIFIED source from /root/reference on top of the restated
tensorly leaves; tests/test_oracle.py checks this module against them.  For X
with <= 3 modes the rank-1 step is the leading singular pair (pinned by
mathematics).  For X with >= 4 modes: *parity unpinned* against real tensorly
(see _cp.py).

The arithmetic follows the reference step for step (same operand order, fp64
factors whatever the dtype of X, in-place deflation in X's own dtype) so that
trip counts agree; only the Python-level loops of missingvals.py are written as
whole-array expressions.
"""

from __future__ import annotations

import os
import sys
from functools import reduce

import numpy as np

_STANDIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tensorly_standin")


def _parafac():
    """Resolve the restated ``parafac`` without leaving the stand-in on
    sys.path for anyone else."""
    try:
        from tensorly.decomposition._cp import parafac  # stand-in (or real) already importable
        return parafac
    except ImportError:
        sys.path.insert(0, _STANDIN)
        try:
            from tensorly.decomposition._cp import parafac
        finally:
            sys.path.remove(_STANDIN)
        return parafac


# --------------------------------------------------------------------------
# leaves
# --------------------------------------------------------------------------

def contract(X, u, miss=None, n_total=None):
    """Z = X x_1 u.  Dense: tpls.py:83.  Masked: missingvals.py:7-20 --
    per column, the dot over observed samples, divided by how many were
    observed, times the total sample count; 0 where nothing was observed.
    ``n_total`` overrides the sample count (virtual shards)."""
    if miss is None:
        return np.einsum("i...,i...->...", X, u)
    n = X.shape[0]
    flat = X.reshape(n, -1)
    obs = ~miss.reshape(n, -1)
    dots = np.where(obs, flat, 0.0).T @ u
    cnt = obs.sum(axis=0)
    tot = n if n_total is None else n_total
    with np.errstate(invalid="ignore", divide="ignore"):
        z = np.where(cnt > 0, dots / cnt * tot, 0.0)
    return z.reshape(X.shape[1:])


def kron_weights(vecs):
    """kron(w2, w3, ...): first vector slowest, matching X.reshape(N, -1)."""
    return reduce(np.kron, vecs)


def project(X, vecs, miss=None):
    """t = X x_2 w2 x_3 w3 ...  Dense: successive mode products, lowest mode
    first (tpls.py:97-99).  Masked: missingvals.py:23-38 -- per row, the dot
    over observed entries, divided by their count, times P (NaN for an
    all-missing row, as in the reference)."""
    if miss is None:
        out = X
        for v in vecs:
            out = np.tensordot(out, v, axes=([1], [0]))
        return out
    n = X.shape[0]
    flat = X.reshape(n, -1)
    obs = ~miss.reshape(n, -1)
    wk = kron_weights(vecs)
    dots = np.where(obs, flat, 0.0) @ wk
    cnt = obs.sum(axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        return dots / cnt * wk.shape[0]


def rank_r_tensor(factors):
    """sum_r a_r o b_r o ... as a dense tensor (util.py:18-20)."""
    rest = np.ones((1, factors[0].shape[1]))
    for f in factors[1:]:
        rest = (rest[:, None, :] * f[None, :, :]).reshape(-1, f.shape[1])
    return (factors[0] @ rest.T).reshape([f.shape[0] for f in factors])


def r2(A, Ahat):
    """1 - ||(Ahat - A) on finite A||^2 / ||A on finite A||^2  (util.py:7-15)."""
    if Ahat.ndim == 2 and A.ndim == 1:
        A = A.reshape(-1, 1)
    assert A.shape == Ahat.shape
    fin = np.isfinite(A)
    a0 = np.nan_to_num(A)
    return 1 - np.linalg.norm(Ahat * fin - a0) ** 2.0 / np.linalg.norm(a0) ** 2.0


def rank1_vectors(Z, tol):
    """Unit weight vectors of the covariance tensor (tpls.py:84-88)."""
    if Z.ndim == 1:
        return [Z / np.linalg.norm(Z)]
    facs = _parafac()(Z, 1, tol=tol, init="svd", normalize_factors=True)[1]
    return [f.reshape(-1) for f in facs]


# --------------------------------------------------------------------------
# fit
# --------------------------------------------------------------------------

def _centre(Xs, Y):
    """tpls.py:44-71 / cmtf.py:44-83: NaN masks, nanmean over samples, centred
    copies (the caller's arrays are never touched)."""
    if Y.ndim == 1:
        Y = Y.reshape(-1, 1)
    assert Y.ndim == 2
    for X in Xs:
        assert X.shape[0] == Y.shape[0]
    miss = [np.isnan(X) for X in Xs]
    has_miss = [bool(m.any()) for m in miss]
    with np.errstate(invalid="ignore"), _quiet():
        x_mean = [np.nanmean(X, axis=0) for X in Xs]
        y_mean = np.nanmean(Y, axis=0)
    return [X - m for X, m in zip(Xs, x_mean)], Y - y_mean, x_mean, y_mean, miss, has_miss


class _quiet:
    def __enter__(self):
        import warnings
        self._c = warnings.catch_warnings()
        self._c.__enter__()
        warnings.simplefilter("ignore")

    def __exit__(self, *a):
        return self._c.__exit__(*a)


def fit(Xs, Y, n_components, tol=1e-8, max_iter=100, r2_mode="reference"):
    """NIPALS tensor-PLS over a list of coupled tensors sharing mode 0.

    Returns a dict: T (N,R) shared scores; W[l][k] (dim_{k+1}, R) loadings of
    tensor l; U (N,R), Q (M,R); coef (R,R) upper-triangular; R2X[l] (R,), R2Y
    (R,); X_mean[l], Y_mean; trips (R,) inner iterations actually taken.

    ``r2_mode="reference"`` evaluates R2X/R2Y the way the reference does
    (dense reconstruction, re-projection of the training set: tpls.py:115-120);
    ``"residual"`` uses the deflated residuals (identical to <= 4e-16,
    SURVEY.md §0.4) and is what a sensible CPU port would do.
    """
    single = not isinstance(Xs, list)
    if single:
        Xs = [Xs]
    R = n_components
    Xc, Yc, x_mean, y_mean, miss, has_miss = _centre(Xs, Y)
    X0 = [x.copy() for x in Xc] if r2_mode == "reference" else None
    Y0 = Yc.copy()
    sst_x = [np.linalg.norm(np.nan_to_num(x)) ** 2.0 for x in Xc]
    N, M = Yc.shape
    T = np.zeros((N, R))
    U = np.zeros((N, R))
    Q = np.zeros((M, R))
    W = [[np.zeros((d, R)) for d in X.shape[1:]] for X in Xs]
    coef = np.zeros((R, R))
    R2X = [np.zeros(R) for _ in Xs]
    R2Y = np.zeros(R)
    trips = np.zeros(R, dtype=np.int64)

    for a in range(R):
        old_u = np.full(N, np.inf)
        U[:, a] = Yc[:, 0]
        for it in range(max_iter):
            trips[a] = it + 1
            ts = []
            for l, X in enumerate(Xc):
                Z = contract(X, U[:, a], miss[l] if has_miss[l] else None)
                for k, v in enumerate(rank1_vectors(Z, tol)):
                    W[l][k][:, a] = v
            for l, X in enumerate(Xc):
                ts.append(project(X, [w[:, a] for w in W[l]], miss[l] if has_miss[l] else None))
            T[:, a] = np.average(ts, axis=0)  # cmtf.py:120 (exact for one tensor)
            Q[:, a] = Yc.T @ T[:, a]
            Q[:, a] /= np.linalg.norm(Q[:, a])
            U[:, a] = Yc @ Q[:, a]
            if np.linalg.norm(old_u - U[:, a]) < tol:
                break
            old_u = U[:, a].copy()

        for l, X in enumerate(Xc):
            X -= rank_r_tensor([T[:, [a]]] + [w[:, [a]] for w in W[l]])
        coef[:, a] = np.linalg.lstsq(T, U[:, a], rcond=-1)[0]
        Yc -= T @ coef[:, [a]] @ Q[:, [a]].T

        if r2_mode == "reference":
            for l in range(len(Xc)):
                R2X[l][a] = r2(X0[l], rank_r_tensor([T] + W[l]))
            state = dict(T=T, W=W, U=U, Q=Q, coef=coef, X_mean=x_mean, Y_mean=y_mean,
                         shapes=[X.shape for X in Xs])
            R2Y[a] = r2(Y0, predict(state, Xs) - y_mean)
        else:
            for l, X in enumerate(Xc):
                R2X[l][a] = 1 - np.linalg.norm(np.nan_to_num(X)) ** 2.0 / sst_x[l]
            R2Y[a] = 1 - np.linalg.norm(Yc) ** 2.0 / np.linalg.norm(Y0) ** 2.0

    return dict(T=T, W=W, U=U, Q=Q, coef=coef, R2X=R2X, R2Y=R2Y, X_mean=x_mean,
                Y_mean=y_mean, trips=trips, shapes=[X.shape for X in Xs])


def transform(state, Xs, Y=None):
    """Scores of new data (tpls.py:145-186, cmtf.py:179-231): centre with the
    training mean, then per component project and deflate with the stored
    loadings; coupled tensors average their projections."""
    if not isinstance(Xs, list):
        Xs = [Xs]
    R = state["T"].shape[1]
    Xw = [X.copy() - m for X, m in zip(Xs, state["X_mean"])]
    miss = [np.isnan(X) for X in Xw]
    has = [bool(m.any()) for m in miss]
    S = np.zeros((Xw[0].shape[0], R))
    for a in range(R):
        ts = [project(X, [w[:, a] for w in state["W"][l]], miss[l] if has[l] else None)
              for l, X in enumerate(Xw)]
        S[:, a] = np.average(ts, axis=0)
        for l, X in enumerate(Xw):
            X -= rank_r_tensor([S[:, [a]]] + [w[:, [a]] for w in state["W"][l]])
    if Y is None:
        return S
    Yw = Y.copy().astype(np.float64)
    if Yw.ndim == 1:
        Yw = Yw.reshape(-1, 1)
    Yw -= state["Y_mean"]
    V = np.zeros((Yw.shape[0], R))
    for a in range(R):
        V[:, a] = Yw @ state["Q"][:, a]
        Yw -= S @ state["coef"][:, [a]] @ state["Q"][:, [a]].T
    return S, V


def predict(state, Xs):
    """tpls.py:122-143 / cmtf.py:142-177."""
    return transform(state, Xs) @ state["coef"] @ state["Q"].T + state["Y_mean"]


def q2y_kfold(Xs, Y, n_components, folds):
    """Cross-validated Q2Y for 1..R components, the slow way: for every fold a refit on
    the sliced training rows and a prediction of the held-out rows with the model
    truncated to k components (formula of validate.py:35-37).  Also returns the
    cross-validated scores."""
    if not isinstance(Xs, list):
        Xs = [Xs]
    Y2 = Y.reshape(Y.shape[0], -1)
    R = n_components
    press = np.zeros(R)
    cv = np.zeros((Y2.shape[0], R))
    for held in folds:
        train = np.setdiff1d(np.arange(Y2.shape[0]), held)
        st = fit([x[train] for x in Xs], Y2[train], R, r2_mode="residual")
        S = transform(st, [x[held] for x in Xs])
        cv[held] = S
        for k in range(1, R + 1):
            pred = S[:, :k] @ st["coef"][:k, :k] @ st["Q"][:, :k].T + st["Y_mean"]
            press[k - 1] += np.sum((pred - Y2[held]) ** 2)
    return 1 - press / np.sum(Y2 ** 2), cv


# --------------------------------------------------------------------------
# virtual shards: the multi-GPU collective contract, on the CPU
# --------------------------------------------------------------------------

def fit_sharded(X_shards, Y_shards, n_components, tol=1e-8, max_iter=100, allreduce=None):
    """The same NIPALS loop with every sample-mode reduction written as
    ``allreduce(local partial)``: column sums/counts, Z, q,
    the stop norm, T'T / T'u, and the residual norms.  ``X_shards`` is a list
    (one entry per tensor) of THIS rank's row blocks when ``allreduce`` is a
    real collective, or a list over ranks of such lists when ``allreduce`` is
    None (all ranks simulated in-process, partials summed in rank order).

    Dense or NaN-carrying tensors, fp64.  Returns the rank-local state (T, U
    are this rank's rows) -- or the list of all ranks' states when simulated.
    """
    simulated = allreduce is None
    ranks_X = X_shards if simulated else [X_shards]
    ranks_Y = Y_shards if simulated else [Y_shards]
    ranks_Y = [y.reshape(-1, 1) if y.ndim == 1 else y for y in ranks_Y]
    G = len(ranks_X)
    L = len(ranks_X[0])
    R = n_components

    def ar(parts):
        if simulated:
            tot = parts[0].copy() if hasattr(parts[0], "copy") else parts[0]
            for p in parts[1:]:
                tot = tot + p
            return tot
        return allreduce(np.asarray(parts[0], dtype=np.float64))

    # means over the global sample mode
    flat = [[x.reshape(x.shape[0], -1) for x in xs] for xs in ranks_X]
    obs = [[~np.isnan(f) for f in fs] for fs in flat]
    masked = [bool(ar([np.array([float((~o[l]).any())]) for o in obs])[0] > 0) for l in range(L)]
    cnt = [ar([o[l].sum(axis=0).astype(np.float64) for o in obs]) for l in range(L)]
    sums = [ar([np.where(o[l], f[l], 0.0).sum(axis=0) for o, f in zip(obs, flat)]) for l in range(L)]
    n_tot = ar([np.array([float(y.shape[0])]) for y in ranks_Y])[0]
    with np.errstate(invalid="ignore", divide="ignore"):
        x_mean = [s / c for s, c in zip(sums, cnt)]
    y_mean = ar([y.sum(axis=0) for y in ranks_Y]) / n_tot
    Xc = [[f[l] - x_mean[l] for l in range(L)] for f in flat]
    Yc = [y - y_mean for y in ranks_Y]
    sst_x = [ar([np.array([np.nansum(x[l] ** 2)]) for x in Xc])[0] for l in range(L)]
    sst_y = ar([np.array([np.sum(y ** 2)]) for y in Yc])[0]
    dims = [x.shape[1:] for x in ranks_X[0]]
    M = Yc[0].shape[1]

    T = [np.zeros((y.shape[0], R)) for y in Yc]
    U = [np.zeros((y.shape[0], R)) for y in Yc]
    Q = np.zeros((M, R))
    W = [[np.zeros((d, R)) for d in dm] for dm in dims]
    coef = np.zeros((R, R))
    R2X = [np.zeros(R) for _ in range(L)]
    R2Y = np.zeros(R)
    trips = np.zeros(R, dtype=np.int64)

    for a in range(R):
        old = [np.full(y.shape[0], np.inf) for y in Yc]
        for g in range(G):
            U[g][:, a] = Yc[g][:, 0]
        for it in range(max_iter):
            trips[a] = it + 1
            for l in range(L):
                z = ar([np.where(obs[g][l], Xc[g][l], 0.0).T @ U[g][:, a] for g in range(G)])
                if masked[l]:
                    with np.errstate(invalid="ignore", divide="ignore"):
                        z = np.where(cnt[l] > 0, z / cnt[l] * n_tot, 0.0)
                for k, v in enumerate(rank1_vectors(z.reshape(dims[l]), tol)):
                    W[l][k][:, a] = v
            for g in range(G):
                ts = []
                for l in range(L):
                    wk = kron_weights([w[:, a] for w in W[l]])
                    d = np.where(obs[g][l], Xc[g][l], 0.0) @ wk
                    if masked[l]:
                        with np.errstate(invalid="ignore", divide="ignore"):
                            d = d / obs[g][l].sum(axis=1) * wk.shape[0]
                    ts.append(d)
                T[g][:, a] = np.average(ts, axis=0)
            q = ar([Yc[g].T @ T[g][:, a] for g in range(G)])
            Q[:, a] = q / np.linalg.norm(q)
            for g in range(G):
                U[g][:, a] = Yc[g] @ Q[:, a]
            with np.errstate(invalid="ignore"):
                d2 = ar([np.array([np.sum((old[g] - U[g][:, a]) ** 2)]) for g in range(G)])[0]
            if np.sqrt(d2) < tol:
                break
            old = [U[g][:, a].copy() for g in range(G)]

        for g in range(G):
            for l in range(L):
                wk = kron_weights([w[:, a] for w in W[l]])
                Xc[g][l] -= np.outer(T[g][:, a], wk)
        k = a + 1
        gram = ar([T[g][:, :k].T @ T[g][:, :k] for g in range(G)])
        rhs = ar([T[g][:, :k].T @ U[g][:, a] for g in range(G)])
        coef[:k, a] = np.linalg.solve(gram, rhs)
        for g in range(G):
            Yc[g] -= T[g] @ coef[:, [a]] @ Q[:, [a]].T
        for l in range(L):
            R2X[l][a] = 1 - ar([np.array([np.nansum(Xc[g][l] ** 2)]) for g in range(G)])[0] / sst_x[l]
        R2Y[a] = 1 - ar([np.array([np.sum(Yc[g] ** 2)]) for g in range(G)])[0] / sst_y

    def state(g):
        return dict(T=T[g], W=W, U=U[g], Q=Q, coef=coef, R2X=R2X, R2Y=R2Y,
                    X_mean=[m.reshape(d) for m, d in zip(x_mean, dims)], Y_mean=y_mean, trips=trips)

    return [state(g) for g in range(G)] if simulated else state(0)


# --------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------

def synthetic(dims, n_response, n_latent, error=0.0, seed=215, extra_dims=()):
    """``import_synthetic`` draw order (synthetic.py:59-77): default_rng(seed);
    T ~ N(0,1) (N, L); y_factor (M, L); one factor per remaining mode;
    X = CP + N(0, error); Y = T y_factor' + N(0, error) -- the noise draws are
    consumed even when error == 0; Y is flattened when M == 1.

    ``extra_dims`` appends further coupled tensors built from the SAME T (the
    reference has no coupled generator; tests/test_cmtf.py:33-36 is the nearest
    pattern): after Y, for each extra tensor, its mode factors then its noise.
    Returns (X or [X0, X1, ...], Y, factors).
    """
    rng = np.random.default_rng(seed)
    n = dims[0]
    t = rng.normal(0, 1, size=(n, n_latent))
    yf = rng.normal(0, 1, size=(n_response, n_latent))
    facs = [t] + [rng.normal(0, 1, size=(d, n_latent)) for d in dims[1:]]
    x = rank_r_tensor(facs)
    x += rng.normal(0, error, size=dims)
    y = t @ yf.T
    y += rng.normal(0, error, size=(n, n_response))
    if y.shape[1] == 1:
        y = y.flatten()
    xs = [x]
    for ed in extra_dims:
        f = [t] + [rng.normal(0, 1, size=(d, n_latent)) for d in ed[1:]]
        xe = rank_r_tensor(f)
        xe += rng.normal(0, error, size=ed)
        xs.append(xe)
    return (xs if extra_dims else x), y, dict(x_factors=facs, y_factor=yf)
