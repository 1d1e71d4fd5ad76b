"""GPU: the single operators of include/tpls_b200.h against numpy on the same
seeded inputs (fp64 accumulate: tolerance 1e-12 relative, Frobenius)."""

import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def eng():
    from cmtf_pls_b200._core import get_engine
    return get_engine(0)


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("n,p,dtype", [(1000, 4096, np.float32), (1000, 4096, np.float64), (37, 48, np.float64),
                                       (5000, 24, np.float32), (333, 520, np.float32), (200, 6000, np.float32),
                                       (1, 8, np.float64), (70000, 512, np.float32)])
@pytest.mark.parametrize("masked", [0, 1])
def test_contract_project(eng, n, p, dtype, masked):
    import torch
    rng = np.random.default_rng(n + p)
    X = rng.normal(size=(n, p)).astype(dtype)
    if masked:
        X[rng.random(X.shape) < 0.2] = np.nan
        X[:, p // 2] = np.nan
    u = rng.normal(size=n)
    w = rng.normal(size=p)
    Xd, ud, wd = _dev(X), _dev(u), _dev(w)
    z = torch.empty(p, dtype=torch.float64, device="cuda")
    t = torch.empty(n, dtype=torch.float64, device="cuda")
    code = 0 if dtype == np.float32 else 1
    lib = eng.lib
    eng._ck(lib.tpls_op_contract(eng.h, Xd.data_ptr(), code, n, p, ud.data_ptr(), masked, z.data_ptr(), None, 1))
    eng._ck(lib.tpls_op_project(eng.h, Xd.data_ptr(), code, n, p, wd.data_ptr(), masked, t.data_ptr(), None, 1))
    X64 = X.astype(np.float64)
    if masked:
        obs = ~np.isnan(X64)
        x0 = np.where(obs, X64, 0.0)
        cnt = obs.sum(0)
        with np.errstate(all="ignore"):
            z_ref = np.where(cnt > 0, (x0.T @ u) / cnt * n, 0.0)
            t_ref = (x0 @ w) / obs.sum(1) * p
    else:
        z_ref, t_ref = X64.T @ u, X64 @ w
    assert _rel(z.cpu().numpy(), z_ref) < 1e-12
    tt = t.cpu().numpy()
    assert np.array_equal(np.isnan(tt), np.isnan(t_ref))
    good = ~np.isnan(t_ref)
    assert _rel(tt[good], t_ref[good]) < 1e-12


@pytest.mark.parametrize("n,p,dtype", [(800, 4096, np.float32), (800, 2048, np.float64), (50, 48, np.float64)])
@pytest.mark.parametrize("masked", [0, 1])
def test_deflate_contract(eng, n, p, dtype, masked):
    import torch
    rng = np.random.default_rng(7)
    X = rng.normal(size=(n, p)).astype(dtype)
    if masked:
        X[rng.random(X.shape) < 0.2] = np.nan
    t, u, w = rng.normal(size=n), rng.normal(size=n), rng.normal(size=p) / np.sqrt(p)
    Xd = _dev(X)
    z = torch.empty(p, dtype=torch.float64, device="cuda")
    ss = torch.empty(1, dtype=torch.float64, device="cuda")
    code = 0 if dtype == np.float32 else 1
    td, wd, ud = _dev(t), _dev(w), _dev(u)   # keep the device copies alive across the call
    eng._ck(eng.lib.tpls_op_deflate_contract(eng.h, Xd.data_ptr(), code, n, p, td.data_ptr(), wd.data_ptr(),
                                             ud.data_ptr(), masked, z.data_ptr(), ss.data_ptr(), None, 1))
    Xn = (X.astype(np.float64) - np.outer(t, w)).astype(dtype)      # numpy's in-place `X -= outer` rounding
    got = Xd.cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(Xn))
    # fused multiply-add vs numpy's multiply-then-subtract: within 2 ulp of the operands' magnitude
    scale = np.abs(X.astype(np.float64)) + np.abs(np.outer(t, w))
    worst = float(np.nanmax(np.abs(got.astype(np.float64) - Xn.astype(np.float64)) / scale))
    assert worst <= 2 * float(np.finfo(dtype).eps), worst
    g64 = np.nan_to_num(got.astype(np.float64))
    assert _rel(z.cpu().numpy(), g64.T @ u) < 1e-12
    assert abs(ss.item() - np.sum(g64 ** 2)) / np.sum(g64 ** 2) < 1e-12


@pytest.mark.parametrize("dims", [(24,), (64, 64), (38, 65), (65, 38), (32, 16), (8, 6, 4), (32, 16, 8), (5, 4, 3, 3),
                                  (100, 4, 3), (2, 2), (1, 7), (120, 90), (4, 3, 3, 2, 2), (3, 3, 2, 2, 2, 2), (3, 2, 2, 2, 2, 2, 2),
                                  (16, 8, 6, 4), (40, 33, 5)])
def test_rank1_matches_restated_parafac(eng, dims):
    import torch
    from oracle import tpls_oracle as orc
    rng = np.random.default_rng(sum(dims))
    # a covariance-like tensor: dominant rank-1 part + structured noise
    vs = [rng.normal(size=d) for d in dims]
    Z = 3.0 * orc.rank_r_tensor([v[:, None] for v in vs]) if len(dims) > 1 else 3.0 * vs[0]
    Z = Z + rng.normal(size=dims)
    ref = orc.rank1_vectors(Z, 1e-8)
    p = int(np.prod(dims))
    zd = _dev(Z.reshape(-1))
    w = torch.zeros(sum(dims), dtype=torch.float64, device="cuda")
    wk = torch.zeros(p, dtype=torch.float64, device="cuda")
    sweeps = C.c_int(0)
    darr = (C.c_int * len(dims))(*dims)
    eng._ck(eng.lib.tpls_op_rank1(eng.h, zd.data_ptr(), len(dims), darr, 1e-8, 0, w.data_ptr(), wk.data_ptr(),
                                  C.byref(sweeps), None, 1))
    got = np.split(w.cpu().numpy(), np.cumsum(dims)[:-1])
    for g, r in zip(got, ref):
        assert _rel(g, r) < 1e-10, dims           # same signs too (largest-|entry| convention)
    assert _rel(wk.cpu().numpy(), orc.kron_weights(ref)) < 1e-10
    if len(dims) >= 2:
        import tensorly.decomposition._cp as cp
        assert sweeps.value == cp.last_sweeps[-1]
