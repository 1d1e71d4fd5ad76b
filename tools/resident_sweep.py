"""Resident trip loop against the streaming kernels over the working-set size (one GPU): where the default switch
(TPLS_RESIDENT_MB) should sit.  Prints ms per fit and us per trip for both."""
import os
import sys

sys.path.insert(0, ".")
import numpy as np
import torch
from cmtf_pls_b200 import tPLS
from oracle import tpls_oracle as orc


def timed(est, X, Y, reps=4):
    est.fit(X, Y)
    ms = []
    for _ in range(reps):
        est.fit(X, Y)
        ms.append(est.stats_["fit_ms"])
    return min(ms), int(est.n_iter_.sum()), est.stats_["resident_loops"]


for dtype in (np.float64, np.float32):
    for n in (2500, 5000, 10000, 20000, 40000):
        X, Y, _ = orc.synthetic((n, 32, 32), 4, 5, error=0.5, seed=3)
        X = torch.from_numpy(X.astype(dtype)).cuda()
        Y = torch.from_numpy(Y).cuda()
        mb = X.numel() * X.element_size() / 1e6
        out = {}
        for mode in ("1", "0"):
            os.environ["TPLS_RESIDENT"] = mode
            out[mode] = timed(tPLS(3), X, Y)
        (a, ta, ra), (b, tb, rb) = out["1"], out["0"]
        print(f"{np.dtype(dtype).name} {n:6d} rows {mb:7.1f} MB  resident {a:7.3f} ms ({a / ta * 1e3:6.1f} us/trip, {ra} loops)   "
              f"streaming {b:7.3f} ms ({b / tb * 1e3:6.1f} us/trip)   trips {ta}/{tb}", flush=True)
        del X, Y
