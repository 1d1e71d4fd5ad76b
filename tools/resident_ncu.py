"""Two fits of the BASELINE configs[1] shape through the resident trip loop (for an ncu capture of one of its launches)."""
import sys

sys.path.insert(0, ".")
import torch
from cmtf_pls_b200 import ctPLS
from oracle import tpls_oracle as orc

Xs, Y, _ = orc.synthetic((10000, 32, 16), 4, 5, error=0.5, seed=215, extra_dims=[(10000, 24)])
Xs = [torch.from_numpy(x).cuda() for x in Xs]
Y = torch.from_numpy(Y).cuda()
est = ctPLS(5)
for _ in range(2):
    est.fit(Xs, Y)
print("trips", est.n_iter_.tolist(), "resident loops", est.stats_["resident_loops"], "fit ms", est.stats_["fit_ms"])
