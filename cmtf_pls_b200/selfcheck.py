"""Parity self-check against the committed golden vectors (``tests/golden/*.npz``: inputs and fitted state minted
by running the reference's unmodified source, ``oracle/make_golden.py``).  No oracle code is involved -- the files
hold the answers -- so the check can run anywhere the estimators run, in particular inside ``bench.py`` on N GPUs
before the timed region (sample-sharded fit vs. the single-process reference result).
"""

from __future__ import annotations

import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def golden_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_golden(case):
    d = np.load(os.path.join(GOLDEN, case + ".npz"))
    g = {k: d[k] for k in d.files}
    L = int(g["n_tensors"])
    g["Xs"] = [g[f"X{l}"] for l in range(L)]
    g["W"] = []
    for l in range(L):
        k = 1
        ws = []
        while f"X{l}_factor{k}" in g:
            ws.append(g[f"X{l}_factor{k}"])
            k += 1
        g["W"].append(ws)
    g["T"] = g["X0_factor0"]
    g["R2X"] = [g[f"R2X{l}"] for l in range(L)]
    g["X_mean"] = [g[f"X{l}_mean"] for l in range(L)]
    if "Xnew0" in g:
        g["Xsnew"] = [g[f"Xnew{l}"] for l in range(L) if f"Xnew{l}" in g]
    return g


def col_err(a, b):
    """max over columns of ||a_j - b_j|| / ||b_j|| (zero columns compared absolutely)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.ndim == 1:
        a, b = a[:, None], b[:, None]
    a = a.reshape(a.shape[0], -1)
    b = b.reshape(b.shape[0], -1)
    worst = 0.0
    for j in range(b.shape[1]):
        nb = np.linalg.norm(b[:, j])
        e = np.linalg.norm(a[:, j] - b[:, j])
        worst = max(worst, e / nb if nb > 0 else e)
    return worst


def aligned_errors(got, ref):
    """Compare two fitted states after per-component sign alignment.

    A tensor-PLS component is defined up to sign flips of its loading vectors: flipping one loading of tensor l
    flips that tensor's projection, which can flip T, q, u and the matching row/column of coef.  Each loading
    column is aligned to the reference by the sign of their inner product; the scores must agree AS THEY ARE up
    to the net sign of the component (a wrong overall sign shows up as an error of ~2)."""
    errs = {}
    R = ref["T"].shape[1]
    for l, ws in enumerate(ref["W"]):
        for k, w_ref in enumerate(ws):
            w = np.array(got["W"][l][k], dtype=np.float64)
            for a in range(R):
                if np.dot(w[:, a], w_ref[:, a]) < 0:
                    w[:, a] = -w[:, a]
            errs[f"W{l}.{k}"] = col_err(w, w_ref)
    s = np.sign(np.sum(np.asarray(got["T"]) * ref["T"], axis=0))   # net sign per component, from the scores
    s[s == 0] = 1
    errs["T"] = col_err(np.asarray(got["T"]) * s, ref["T"])
    errs["U"] = col_err(np.asarray(got["U"]) * s, ref["U"])
    errs["Q"] = col_err(np.asarray(got["Q"]) * s, ref["Q"])
    c = np.asarray(got["coef"]) * s[:, None] * s[None, :]
    errs["coef"] = np.linalg.norm(c - ref["coef"]) / np.linalg.norm(ref["coef"])
    errs["R2Y"] = float(np.max(np.abs(np.asarray(got["R2Y"]) - ref["R2Y"])))
    for l in range(len(ref["R2X"])):
        errs[f"R2X{l}"] = float(np.max(np.abs(np.asarray(got["R2X"][l]) - ref["R2X"][l])))
    return errs


def sharded_golden_check(cases, device, process_group=None, rank=0, world=1, gather=None):
    """Fit every golden case with its rows sharded over ``world`` ranks and compare with the stored reference state.

    ``gather(local_rows, n_total, lo, hi)`` must return the full-length array on every rank (bench.py passes an
    all-reduce of a zero-padded copy); with ``world == 1`` nothing is gathered.  Returns
    ``{"cases": [...], "max_err": float, "trips_equal": bool, "worst": "<case>:<what>"}`` -- identical on all ranks.
    """
    from . import ctPLS
    from .sharding import shard_rows, row_block
    worst, worst_at, trips_equal = 0.0, "", True
    for case in cases:
        g = load_golden(case)
        R = int(g["n_components"])
        n = g["Y"].shape[0]
        est = ctPLS(R, device=device, process_group=process_group)
        lo, hi = row_block(n, rank, world)
        import contextlib
        import io
        kw = {"max_iter": 4} if "maxiter" in case else {}     # that case was minted with max_iter=4 (oracle/make_golden.py)
        with contextlib.redirect_stdout(io.StringIO()):       # the "missing values" notice of the reference
            est.fit(shard_rows(list(g["Xs"]), rank, world), shard_rows(g["Y"], rank, world), **kw)
        T, U = est.factor_T, est.Y_factors[0]
        if world > 1:
            T, U = gather(T, n, lo, hi), gather(U, n, lo, hi)
        got = dict(T=T, U=U, Q=est.Y_factors[1], coef=est.coef_, R2Y=est.R2Y, R2X=est.R2Xs,
                   W=[f[1:] for f in est.Xs_factors])
        if est.n_iter_.tolist() != g["trips"].tolist():
            trips_equal = False
        for k, e in aligned_errors(got, g).items():
            if not (e <= worst):      # NaN counts as the worst
                worst, worst_at = (float(e) if e == e else float("inf")), f"{case}:{k}"
    return {"cases": list(cases), "max_err": worst, "trips_equal": trips_equal, "worst": worst_at, "world": world}
