"""Tiny end-to-end runs of every code path (written for compute-sanitizer, which is closed on this pool;
   kept as a quick plain smoke run: python tools/sanitize_case.py).  Under a sanitizer, one tool per call:
   compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from _util import load_golden, aligned_errors  # noqa: E402
from cmtf_pls_b200 import tPLS, ctPLS  # noqa: E402


def state(est, coupled):
    if coupled:
        return dict(T=est.factor_T, W=[f[1:] for f in est.Xs_factors], U=est.Y_factors[0], Q=est.Y_factors[1],
                    coef=est.coef_, R2X=est.R2Xs, R2Y=est.R2Y)
    return dict(T=est.X_factors[0], W=[est.X_factors[1:]], U=est.Y_factors[0], Q=est.Y_factors[1], coef=est.coef_,
                R2X=[est.R2X], R2Y=est.R2Y)


for case in ("c1_20x8x6_m1_r3", "t3_miss_70x12x8_m4_r4", "ct_90x32x16_90x24_m4_r5", "t4_60x8x6x4_m3_r4"):
    g = load_golden(case)
    coupled = bool(g["coupled"])
    for alg in ("stream", "covariance"):
        est = (ctPLS if coupled else tPLS)(int(g["n_components"]), algorithm=alg)
        est.fit([x.copy() for x in g["Xs"]] if coupled else g["Xs"][0].copy(), g["Y"].copy())
        worst = max(aligned_errors(state(est, coupled), g).values())
        assert worst < 1e-8, (case, alg, worst)
        if "Xsnew" in g:
            est.transform([x.copy() for x in g["Xsnew"]] if coupled else g["Xsnew"][0].copy())
        print(case, alg, "ok", est.n_iter_.tolist(), flush=True)
# one larger shape that exercises the FULL fast path and several tiles per CTA
rng = np.random.default_rng(0)
X = rng.normal(size=(3000, 64, 64)).astype(np.float32)
Y = rng.normal(size=(3000, 4))
tPLS(2).fit(X, Y)
print("big ok")
