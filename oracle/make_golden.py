"""TEST INFRASTRUCTURE ONLY -- mint tests/golden/*.npz.

Runs the reference's UNMODIFIED source (``/root/reference/cmtf_pls``) on top of
the restated tensorly leaves (oracle/tensorly_standin) on seeded inputs, and
stores inputs + every fitted attribute of SURVEY.md §8 row a15 + the per-
component trip counts (parsed from the reference's own ``verbose`` output,
tpls.py:104-105).  /root/reference exists only in the build container, so this
script runs there; the fixtures it writes are committed and travel.

    python oracle/make_golden.py            # rewrites tests/golden/

The reference has no golden vectors of its own; these are the
known-answer files the oracle restatement and the CUDA path are held to.
"""

import contextlib
import io
import os
import re
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def _import_reference():
    if not os.path.isdir(REFERENCE):
        raise SystemExit("make_golden.py needs /root/reference (build container only)")
    sys.path.insert(0, os.path.join(HERE, "tensorly_standin"))
    sys.path.insert(0, REFERENCE)
    from cmtf_pls.tpls import tPLS
    from cmtf_pls.cmtf import ctPLS
    from cmtf_pls.synthetic import import_synthetic
    return tPLS, ctPLS, import_synthetic


def _fit_capture(est, X, Y, max_iter=100):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        est.fit(X, Y, verbose=1, max_iter=max_iter)
    trips = np.full(est.n_components, max_iter, dtype=np.int64)
    for m in re.finditer(r"Comp (\d+): converged after (\d+) iterations", buf.getvalue()):
        trips[int(m.group(1))] = int(m.group(2)) + 1
    return trips


def _pack(case, Xs, Y, est, trips, coupled, extra=None):
    d = {"coupled": np.array(coupled), "n_tensors": np.array(len(Xs)), "Y": Y,
         "n_components": np.array(est.n_components), "trips": trips,
         "coef": est.coef_, "R2Y": est.R2Y, "Y_mean": est.Y_mean,
         "U": est.Y_factors[0], "Q": est.Y_factors[1]}
    for l, X in enumerate(Xs):
        d[f"X{l}"] = X
        facs = est.Xs_factors[l] if coupled else est.X_factors
        d[f"X{l}_mean"] = est.Xs_mean[l] if coupled else est.X_mean
        d[f"R2X{l}"] = est.R2Xs[l] if coupled else est.R2X
        for k, f in enumerate(facs):
            d[f"X{l}_factor{k}"] = f
    if extra:
        d.update(extra)
    path = os.path.join(OUT, case + ".npz")
    np.savez_compressed(path, **d)
    print(f"{case:28s} trips={trips.tolist()}  {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    tPLS, ctPLS, import_synthetic = _import_reference()
    os.makedirs(OUT, exist_ok=True)

    def single(case, X, Y, R, Xnew=None, Ynew=None, max_iter=100):
        est = tPLS(R)
        trips = _fit_capture(est, X, Y, max_iter)
        extra = {}
        if Xnew is not None:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                extra["Xnew0"] = Xnew
                extra["predict_new"] = est.predict(Xnew)
                if Ynew is not None:
                    extra["Ynew"] = Ynew
                    s, v = est.transform(Xnew, Ynew)
                    extra["transform_new_X"], extra["transform_new_Y"] = s, v
                else:
                    extra["transform_new_X"] = est.transform(Xnew)
        _pack(case, [X], Y, est, trips, False, extra)

    def coupled(case, Xs, Y, R, Xsnew=None, Ynew=None):
        est = ctPLS(R)
        trips = _fit_capture(est, Xs, Y)
        extra = {}
        if Xsnew is not None:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                for l, x in enumerate(Xsnew):
                    extra[f"Xnew{l}"] = x
                extra["predict_new"] = est.predict(Xsnew)
                extra["Ynew"] = Ynew
                s, v = est.transform(Xsnew, Ynew)
                extra["transform_new_X"], extra["transform_new_Y"] = s, v
        _pack(case, Xs, Y, est, trips, True, extra)

    # c1: BASELINE.json configs[0] -- 20x8x6, one response, 3 components
    X, Y, _ = import_synthetic((20, 8, 6), 1, 3, error=0.1, seed=215)
    single("c1_20x8x6_m1_r3", X, Y, 3)

    # 3-way, several responses (matrix Z: pinned by mathematics), with new data
    X, Y, _ = import_synthetic((60, 16, 12), 4, 6, error=0.5, seed=215)
    Xn, Yn, _ = import_synthetic((9, 16, 12), 4, 6, error=0.5, seed=7)
    single("t3_60x16x12_m4_r5", X, Y, 5, Xn, Yn)

    # 2-way X (vector Z), M = 1
    X, Y, _ = import_synthetic((50, 12), 1, 4, error=0.3, seed=11)
    single("t2_50x12_m1_r3", X, Y, 3)

    # 2-way X, many responses (the PCA-like case of tests/test_tpls.py:84-95)
    X, _, _ = import_synthetic((40, 30), 4, 8, error=0.0, seed=215)
    single("t2_same_xy_40x30_r4", X, X.copy(), 4)

    # 4-way X (rank-1 CP of a 3-way Z: parity unpinned vs real tensorly)
    X, Y, _ = import_synthetic((60, 8, 6, 4), 3, 5, error=0.5, seed=215)
    Xn, Yn, _ = import_synthetic((7, 8, 6, 4), 3, 5, error=0.5, seed=8)
    single("t4_60x8x6x4_m3_r4", X, Y, 4, Xn, Yn)

    # 5-way X
    X, Y, _ = import_synthetic((30, 5, 4, 3, 3), 2, 3, error=0.3, seed=3)
    single("t5_30x5x4x3x3_m2_r3", X, Y, 3)

    # more components than latent structure
    X, Y, _ = import_synthetic((40, 10, 6), 3, 2, error=0.05, seed=5)
    single("t3_overfit_40x10x6_m3_r4", X, Y, 4)

    # constant slab -> zero loading row (tests/test_tpls.py:98-104)
    X, Y, _ = import_synthetic((40, 9, 7), 4, 5, error=0.2, seed=215)
    X[:, 0, :] = 1
    single("t3_constslab_40x9x7_m4_r4", X, Y, 4)

    # fp32 storage
    X, Y, _ = import_synthetic((80, 16, 8), 4, 6, error=0.5, seed=21)
    single("t3_f32_80x16x8_m4_r4", X.astype(np.float32), Y, 4)

    # missing values: 20 % NaN, incl. one all-NaN column; masked new data too
    rng = np.random.default_rng(99)
    X, Y, _ = import_synthetic((70, 12, 8), 4, 5, error=0.5, seed=31)
    X[rng.random(X.shape) < 0.2] = np.nan
    X[:, 3, 5] = np.nan
    Xn, Yn, _ = import_synthetic((6, 12, 8), 4, 5, error=0.5, seed=32)
    Xn[rng.random(Xn.shape) < 0.2] = np.nan
    single("t3_miss_70x12x8_m4_r4", X, Y, 4, Xn, Yn)

    # missing values, 4-way, fp32
    X, Y, _ = import_synthetic((50, 6, 5, 4), 3, 4, error=0.3, seed=41)
    X[rng.random(X.shape) < 0.15] = np.nan
    single("t4_miss_f32_50x6x5x4_m3_r3", X.astype(np.float32), Y, 3)

    # coupled: 3-way + matrix sharing T (down-scaled BASELINE configs[1])
    gen = np.random.default_rng(215)
    T = gen.normal(size=(90, 8))
    yf = gen.normal(size=(4, 8))
    A, B, C = gen.normal(size=(32, 8)), gen.normal(size=(16, 8)), gen.normal(size=(24, 8))
    X0 = np.einsum("ir,jr,kr->ijk", T, A, B) + gen.normal(0, 0.5, size=(90, 32, 16))
    X1 = T @ C.T + gen.normal(0, 0.5, size=(90, 24))
    Y = T @ yf.T + gen.normal(0, 0.5, size=(90, 4))
    Tn = gen.normal(size=(5, 8))
    Xn0 = np.einsum("ir,jr,kr->ijk", Tn, A, B) + gen.normal(0, 0.5, size=(5, 32, 16))
    Xn1 = Tn @ C.T + gen.normal(0, 0.5, size=(5, 24))
    Yn = Tn @ yf.T
    coupled("ct_90x32x16_90x24_m4_r5", [X0, X1], Y, 5, [Xn0, Xn1], Yn)

    # coupled: 4-way + 3-way + matrix, one of them with NaNs (cmtf.py:93-96,112-117)
    dims = [(30, 6, 5, 4), (30, 5, 4), (30, 7)]
    Xs = [gen.random(d) for d in dims]
    Xs[1][gen.random(dims[1]) < 0.1] = np.nan
    Y = gen.random((30, 5))
    coupled("ct_mixed_miss_30_m5_r3", Xs, Y, 3)

    # ctPLS([X]) next to tPLS(X) (tests/test_cmtf.py:8-15)
    X, Y, _ = import_synthetic((30, 7, 6), 3, 4, error=0.4, seed=77)
    coupled("ct_single_30x7x6_m3_r3", [X], Y, 3)
    single("t3_single_30x7x6_m3_r3", X, Y, 3)

    # does not converge within max_iter (silent in the reference, tpls.py:79-107)
    X = gen.random((20, 6, 5, 4))
    Y = gen.random((20, 5))
    single("t4_maxiter_20x6x5x4_m5_r2", X, Y, 2, max_iter=4)

    # 6-way and 7-way X (5- and 6-way Z through the HOSVD start + ALS sweeps): own generator, so that the cases
    # above regenerate bit-identically
    X, Y, _ = import_synthetic((24, 4, 3, 3, 2, 2), 3, 3, error=0.3, seed=61)
    single("t6_24x4x3x3x2x2_m3_r2", X, Y, 2)
    X, Y, _ = import_synthetic((20, 3, 3, 2, 2, 2, 2), 2, 3, error=0.3, seed=71)
    Xn, Yn, _ = import_synthetic((5, 3, 3, 2, 2, 2, 2), 2, 3, error=0.3, seed=72)
    single("t7_20x3x3x2x2x2x2_m2_r2", X, Y, 2, Xn, Yn)


if __name__ == "__main__":
    main()
