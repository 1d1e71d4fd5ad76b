"""Shared helpers for the parity tests: golden loading, per-component sign
alignment and column-wise Frobenius-relative error."""

import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


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
    """max over columns of ||a_j - b_j|| / ||b_j|| (zero columns compared
    absolutely)."""
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

    A tensor-PLS component is defined up to sign flips of its loading vectors:
    flipping one loading of tensor l flips that tensor's projection, which can
    flip T, q, u and the matching row/column of coef.  We align each loading
    column to the reference by the sign of their inner product and require the
    scores to agree AS THEY ARE (a wrong overall sign shows up as error ~2)."""
    errs = {}
    R = ref["T"].shape[1]
    for l, ws in enumerate(ref["W"]):
        for k, w_ref in enumerate(ws):
            w = np.array(got["W"][l][k], dtype=np.float64)
            for a in range(R):
                if np.dot(w[:, a], w_ref[:, a]) < 0:
                    w[:, a] = -w[:, a]
            errs[f"W{l}.{k}"] = col_err(w, w_ref)
    # net sign per component, taken from the scores
    s = np.sign(np.sum(np.asarray(got["T"]) * ref["T"], axis=0))
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
