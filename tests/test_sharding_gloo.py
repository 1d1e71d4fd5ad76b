"""CPU, world_size 2 and 3 over gloo: the multi-GPU collective contract
 executed by real processes -- every sample-mode reduction of the
NIPALS loop is a torch.distributed all-reduce of a small replicated quantity --
must reproduce the single-process golden fit, trip counts included."""

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from _util import load_golden
    from cmtf_pls_b200.sharding import shard_rows
    from oracle import tpls_oracle as orc
    g = load_golden(case)
    Xs = shard_rows([x for x in g["Xs"]], rank, world)
    Y = shard_rows(g["Y"], rank, world)
    calls = [0]

    def allreduce(a):
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).copy())
        dist.all_reduce(t)
        calls[0] += 1
        return t.numpy()

    st = orc.fit_sharded(Xs, Y, int(g["n_components"]), allreduce=allreduce)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), T=st["T"], U=st["U"], Q=st["Q"], coef=st["coef"], R2Y=st["R2Y"],
             trips=st["trips"], calls=np.array(calls[0]), **{f"R2X{l}": r for l, r in enumerate(st["R2X"])},
             **{f"W{l}_{k}": w for l, ws in enumerate(st["W"]) for k, w in enumerate(ws)})
    dist.destroy_process_group()


@pytest.mark.parametrize("case,world", [("t3_60x16x12_m4_r5", 2), ("ct_90x32x16_90x24_m4_r5", 2),
                                        ("t3_miss_70x12x8_m4_r4", 3)])
def test_sharded_fit_over_gloo_matches_golden(tmp_path, case, world):
    from _util import load_golden, aligned_errors
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)
    g = load_golden(case)
    parts = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    whole = dict(T=np.concatenate([p["T"] for p in parts]), U=np.concatenate([p["U"] for p in parts]),
                 Q=parts[0]["Q"], coef=parts[0]["coef"], R2Y=parts[0]["R2Y"],
                 R2X=[parts[0][f"R2X{l}"] for l in range(len(g["Xs"]))],
                 W=[[parts[0][f"W{l}_{k}"] for k in range(len(ws))] for l, ws in enumerate(g["W"])])
    assert parts[0]["trips"].tolist() == g["trips"].tolist()
    for k, e in aligned_errors(whole, g).items():
        assert e < 1e-8, (case, k, e)
    # replicated quantities are bit-identical on every rank (they come out of all-reduces)
    for p in parts[1:]:
        assert np.array_equal(p["Q"], parts[0]["Q"]) and np.array_equal(p["coef"], parts[0]["coef"])
    # payload discipline: a fixed number of collectives before the loop, 3 per trip per tensor-set, 3 per component
    assert int(parts[0]["calls"]) > 0


def test_row_block_partitions_exactly():
    from cmtf_pls_b200.sharding import row_block
    for n in (1, 7, 8, 1000, 1_000_000):
        for w in (1, 2, 3, 8):
            blocks = [row_block(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        row_block(10, 3, 3)


# --------------------------------------------------------------------------
# fold-parallel cross-validation sweep (folds are independent units)
# --------------------------------------------------------------------------
def _fold_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cmtf_pls_b200.validate import q2y_sweep, _folds
    from oracle import tpls_oracle as orc
    X, Y, _ = orc.synthetic((40, 8, 6), 3, 3, error=0.4, seed=9)
    fitted = []

    def fit_fold(w):            # CPU stand-in for the device fit: refit on the kept rows, project every row
        keep = w > 0
        st = orc.fit([X[keep]], Y[keep], 3, r2_mode="residual")
        fitted.append(int(np.flatnonzero(~keep)[0]))
        return dict(T=orc.transform(st, [X]), coef=st["coef"], Q=st["Q"], Y_mean=st["Y_mean"])

    folds = _folds(40, 5, 3)
    q, cv = q2y_sweep(X, Y, 3, folds=folds, return_scores=True, fold_group=True, _fit_fold=fit_fold)
    np.savez(os.path.join(out_dir, f"f{rank}.npz"), q=q, cv=cv, n_fits=np.array(len(fitted)))
    dist.destroy_process_group()


def test_fold_parallel_sweep_over_gloo(tmp_path):
    from cmtf_pls_b200.validate import _folds
    from oracle import tpls_oracle as orc
    world = 2
    mp.spawn(_fold_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    X, Y, _ = orc.synthetic((40, 8, 6), 3, 3, error=0.4, seed=9)
    q_ref, cv_ref = orc.q2y_kfold(X, Y, 3, _folds(40, 5, 3))
    parts = [np.load(tmp_path / f"f{r}.npz") for r in range(world)]
    assert [int(p["n_fits"]) for p in parts] == [3, 2]          # folds 0, 2, 4 on rank 0; 1, 3 on rank 1
    for p in parts:                                             # every rank returns the complete result
        assert np.max(np.abs(p["q"] - q_ref)) < 1e-10
        assert np.max(np.abs(p["cv"] - cv_ref)) < 1e-9 * np.max(np.abs(cv_ref))
