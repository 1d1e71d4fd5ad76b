"""GPU, >= 2 devices: sample-sharded fits over NCCL (one process per GPU, launched
with torchrun) must agree with the oracle / golden vectors -- shard-count
invariance of the collective contract.  Skipped on a
single-GPU box; the CPU counterpart is tests/test_sharding_gloo.py."""

import os
import subprocess
import sys

import numpy as np
import pytest

from _util import load_golden, aligned_errors

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _run(case, world, out):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + (os.getpid() % 400)),
           os.path.join(ROOT, "tests", "_multi_worker.py"), case, out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]


def _whole(path, n_tensors, n_modes):
    p = np.load(path)
    return dict(T=p["T"], U=p["U"], Q=p["Q"], coef=p["coef"], R2Y=p["R2Y"], R2X=[p[f"R2X{l}"] for l in range(n_tensors)],
                W=[[p[f"W{l}_{k}"] for k in range(n_modes[l])] for l in range(n_tensors)]), p


@pytest.mark.parametrize("case", ["ct_90x32x16_90x24_m4_r5", "t3_miss_70x12x8_m4_r4", "t4_60x8x6x4_m3_r4",
                                  "t2_same_xy_40x30_r4"])   # the last one has 30 responses: explicit ||du||^2 exchange
def test_sharded_fit_matches_golden(tmp_path, case):
    world = min(_ngpu(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    g = load_golden(case)
    out = str(tmp_path / "out.npz")
    _run(case, world, out)
    whole, p = _whole(out, len(g["Xs"]), [len(w) for w in g["W"]])
    assert p["trips"].tolist() == g["trips"].tolist()
    for k, e in aligned_errors(whole, g).items():
        assert e < 1e-8, (case, k, e)
    assert int(p["collectives"]) >= 2 * int(p["trips"].sum())   # Z and q every trip; the stop test rides on q


@pytest.mark.parametrize("case,tol", [("synthetic_f64", 1e-8), ("synthetic_f32", 1e-4), ("synthetic_miss_f64", 1e-8)])
def test_sharded_fit_matches_oracle_c4_rows(tmp_path, case, tol):
    world = min(_ngpu(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    from oracle import tpls_oracle as orc
    dt = np.float32 if case.endswith("f32") else np.float64
    Xs, Y, _ = orc.synthetic((3000, 64, 64), 4, 6, error=0.7, seed=5, extra_dims=[(3000, 32, 16)])
    Xs = [x.astype(dt) for x in Xs]
    if "miss" in case:
        rng = np.random.default_rng(1)
        Xs[0][rng.random(Xs[0].shape) < 0.15] = np.nan
    ref = orc.fit([x.copy() for x in Xs], Y.copy(), 4, r2_mode="residual")
    out = str(tmp_path / "out.npz")
    _run(case, world, out)
    whole, p = _whole(out, 2, [2, 2])
    assert p["trips"].tolist() == ref["trips"].tolist()
    for k, e in aligned_errors(whole, ref).items():
        assert e < tol, (case, k, e)


def test_fold_parallel_cv_sweep_matches_oracle_refits(tmp_path):
    """the folds of the cross-validation sweep are independent units -- one fold per GPU,
    only the PRESS sums cross devices."""
    world = min(_ngpu(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    from oracle import tpls_oracle as orc
    from cmtf_pls_b200.validate import _folds
    X, Y, _ = orc.synthetic((90, 12, 8), 3, 4, error=0.4, seed=9)
    q_ref, cv_ref = orc.q2y_kfold(X, Y, 4, _folds(90, 5, 3))
    out = str(tmp_path / "cv.npz")
    _run("cvfold", world, out)
    p = np.load(out)
    assert np.max(np.abs(p["q"] - q_ref)) < 1e-8
    s = np.sign(np.sum(p["cv"] * cv_ref, axis=0))
    assert np.max(np.abs(p["cv"] * s - cv_ref)) / np.max(np.abs(cv_ref)) < 1e-8
