"""torchrun worker of tests/test_gpu_multi.py: every rank fits its block of rows
with the NCCL-backed estimator; rank 0 gathers the scores and writes them out."""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    case, out = sys.argv[1], sys.argv[2]
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    from cmtf_pls_b200 import ctPLS
    from cmtf_pls_b200.sharding import shard_rows, row_block
    if case == "cvfold":
        # fold-parallel sweep: every rank holds all rows and fits its own folds, no data-path collective
        from oracle import tpls_oracle as orc
        from cmtf_pls_b200 import q2y_sweep
        from cmtf_pls_b200.validate import _folds
        X, Y, _ = orc.synthetic((90, 12, 8), 3, 4, error=0.4, seed=9)
        q, cv = q2y_sweep(X, Y, 4, folds=_folds(90, 5, 3), return_scores=True, device=local, fold_group=True)
        if rank == 0:
            np.savez(out, q=q, cv=cv)
        dist.destroy_process_group()
        return
    if case.startswith("synthetic"):
        from oracle import tpls_oracle as orc
        dt = np.float32 if case.endswith("f32") else np.float64
        Xs, Y, _ = orc.synthetic((3000, 64, 64), 4, 6, error=0.7, seed=5, extra_dims=[(3000, 32, 16)])
        Xs = [x.astype(dt) for x in Xs]
        if "miss" in case:
            rng = np.random.default_rng(1)
            Xs[0][rng.random(Xs[0].shape) < 0.15] = np.nan
        R = 4
    else:
        from _util import load_golden
        g = load_golden(case)
        Xs, Y, R = g["Xs"], g["Y"], int(g["n_components"])
    n = Y.shape[0]
    est = ctPLS(R, device=local, process_group=True)
    est.fit(shard_rows(list(Xs), rank, world), shard_rows(Y, rank, world))
    # gather the row-sharded scores on rank 0 (plumbing through torch.distributed)
    T = torch.zeros(n, R, dtype=torch.float64, device="cuda")
    U = torch.zeros(n, R, dtype=torch.float64, device="cuda")
    lo, hi = row_block(n, rank, world)
    T[lo:hi] = torch.from_numpy(est.factor_T).cuda()
    U[lo:hi] = torch.from_numpy(est.Y_factors[0]).cuda()
    dist.all_reduce(T)
    dist.all_reduce(U)
    q = torch.from_numpy(est.Y_factors[1]).cuda()
    qs = [torch.empty_like(q) for _ in range(world)]
    dist.all_gather(qs, q)
    if rank == 0:
        assert all(torch.equal(qs[0], x) for x in qs), "replicated Q differs between ranks"
        d = dict(T=T.cpu().numpy(), U=U.cpu().numpy(), Q=est.Y_factors[1], coef=est.coef_, R2Y=est.R2Y,
                 trips=est.n_iter_, collectives=np.array(est.stats_["collectives"]))
        for l in range(len(Xs)):
            d[f"R2X{l}"] = est.R2Xs[l]
            for k, w in enumerate(est.Xs_factors[l][1:]):
                d[f"W{l}_{k}"] = w
        np.savez(out, **d)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
