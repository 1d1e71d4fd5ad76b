"""Per-phase timing of the resident trip loop (TPLS_RESIDENT_STAMPS=1) on BASELINE configs[0] / [1] shapes."""
import os
import sys

os.environ["TPLS_RESIDENT_STAMPS"] = "1"
sys.path.insert(0, ".")
import numpy as np
import torch
from cmtf_pls_b200 import ctPLS, tPLS
from oracle import tpls_oracle as orc


def timed(est, Xs, Y, reps=5):
    est.fit(Xs, Y)
    ms = []
    for _ in range(reps):
        est.fit(Xs, Y)
        ms.append(est.stats_["fit_ms"])
    return min(ms), int(est.n_iter_.sum()), est.stats_["kernel_launches"]


X, Y, _ = orc.synthetic((20, 8, 6), 1, 3, error=0.3, seed=215)
print("configs[0]", timed(tPLS(3), torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda()), flush=True)
Xs, Y, _ = orc.synthetic((10000, 32, 16), 4, 5, error=0.5, seed=215, extra_dims=[(10000, 24)])
Xs = [torch.from_numpy(x).cuda() for x in Xs]
Y = torch.from_numpy(Y).cuda()
print("configs[1]", timed(ctPLS(5), Xs, Y), flush=True)
os.environ["TPLS_RESIDENT"] = "0"
print("configs[1] streaming kernels", timed(ctPLS(5), Xs, Y), flush=True)
