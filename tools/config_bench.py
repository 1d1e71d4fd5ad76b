"""Fit time of every BASELINE.json configuration at FULL size on one B200, both
inner-loop algorithms, beside the CPU oracle port on a bounded sample of the same
shape.  One JSON line per configuration:

    python tools/config_bench.py [--configs 1,2,3,4,5] [--no-cpu] > profiles/rNN_configs.jsonl

`ms_device` is the CUDA-event time of the whole device fit (tpls_stats.fit_ms,
data resident in HBM); `gbs` = B_alg / that (s*N*P*(2*trips + R + 2)
summed over the coupled tensors).  bench.py remains the judged benchmark
(configs[3]); this file documents the other configurations.
"""

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cp_data(n, dims_list, M, L, error, dtype, seed, nan_frac=0.0, device="cuda"):
    """Coupled CP-structured tensors sharing the scores T (synthetic.py:59-77 pattern), drawn on the device."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    T = torch.randn(n, L, generator=g, device=device, dtype=torch.float64)
    yf = torch.randn(M, L, generator=g, device=device, dtype=torch.float64)
    Xs = []
    for dims in dims_list:
        kr = torch.ones(1, L, device=device, dtype=torch.float64)
        for d in dims:
            f = torch.randn(d, L, generator=g, device=device, dtype=torch.float64)
            kr = (kr[:, None, :] * f[None, :, :]).reshape(-1, L)
        X = torch.empty(n, *dims, dtype=dtype, device=device)
        Xf = X.view(n, -1)
        step = 32768
        for r0 in range(0, n, step):
            r1 = min(n, r0 + step)
            blk = (T[r0:r1] @ kr.T).to(dtype)
            blk += error * torch.randn(r1 - r0, kr.shape[0], generator=g, device=device, dtype=dtype)
            if nan_frac:
                blk[torch.rand(r1 - r0, kr.shape[0], generator=g, device=device) < nan_frac] = float("nan")
            Xf[r0:r1] = blk
        Xs.append(X)
    Y = T @ yf.T + error * torch.randn(n, M, generator=g, device=device, dtype=torch.float64)
    return Xs, Y


CONFIGS = {
    1: dict(name="configs[0] single 3-way 20x8x6, Y 20x1, R=3, fp64", n=20, dims=[(8, 6)], M=1, L=3, error=0.1, R=3,
            dtype="float64", cpu_rows=20),
    2: dict(name="configs[1] coupled 10k x 32x16 + 10k x 24, Y 10k x 4, R=5, fp64", n=10_000, dims=[(32, 16), (24,)], M=4,
            L=8, error=0.5, R=5, dtype="float64", cpu_rows=10_000),
    3: dict(name="configs[2] 100k x 64x32, 20% NaN, R=5, fp32", n=100_000, dims=[(64, 32)], M=4, L=8, error=0.5, R=5,
            dtype="float32", nan=0.2, cpu_rows=2_000),
    4: dict(name="configs[3] coupled pair 2 x [1M x 64x64] fp32, R=10", n=1_000_000, dims=[(64, 64), (64, 64)], M=4, L=12,
            error=1.0, R=10, dtype="float32", cpu_rows=0),
    5: dict(name="configs[4] 4-way 200k x 32x16x8 fp32, R=10, 5-fold sweep over 1..10 components", n=200_000,
            dims=[(32, 16, 8)], M=4, L=12, error=1.0, R=10, dtype="float32", cpu_rows=2_000, cv=5),
}


def arg_of(cls, Xs):
    return Xs if cls.__name__ == "ctPLS" else Xs[0]


def timed_fit(cls, Xs, Y, R, algorithm, repeats):
    import torch
    arg = arg_of(cls, Xs)
    est = cls(R, algorithm=algorithm)
    est.fit(arg, Y)                               # warm-up (buffers, first-touch)
    dev_ms, wall = [], []
    for _ in range(repeats):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        est.fit(arg, Y)
        torch.cuda.synchronize()
        wall.append(1e3 * (time.perf_counter() - t0))
        dev_ms.append(est.stats_["fit_ms"])
    return est, float(np.median(dev_ms)), float(np.median(wall))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4,5")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--repeats", type=int, default=3)
    args = ap.parse_args()
    import torch
    from cmtf_pls_b200 import tPLS, ctPLS, q2y_sweep, trim_memory

    real_stdout = sys.stdout
    sys.stdout = sys.stderr          # the estimators print the reference's "X has missing values" notice
    for ci in [int(c) for c in args.configs.split(",")]:
        c = CONFIGS[ci]
        dtype = getattr(torch, c["dtype"])
        Xs, Y = cp_data(c["n"], c["dims"], c["M"], c["L"], c["error"], dtype, 215, c.get("nan", 0.0))
        if c["M"] == 1:
            Y = Y.reshape(-1)
        cls = ctPLS if len(Xs) > 1 else tPLS
        line = {"config": c["name"], "rows": c["n"], "dtype": c["dtype"], "components": c["R"], "gpu": torch.cuda.get_device_name(0)}
        elem = 4 if c["dtype"] == "float32" else 8
        pass_bytes = sum(elem * c["n"] * int(np.prod(d)) for d in c["dims"])
        for alg in ("stream", "covariance"):
            est, ms, wall = timed_fit(cls, Xs, Y, c["R"], alg, args.repeats)
            trips = int(est.n_iter_.sum())
            b_alg = pass_bytes * (2.0 * trips + c["R"] + 2)
            d = {"ms_device": ms, "ms_wall_incl_readback": wall, "trips": est.n_iter_.tolist(),
                 "kernel_launches": int(est.stats_["kernel_launches"])}
            if alg == "stream":
                d["b_alg_gb"] = b_alg / 1e9
                d["gbs"] = b_alg / (ms * 1e-3) / 1e9
            else:
                d["ran_covariance_loop"] = bool(est.stats_["covariance_mode"])
                d["streamed_gb"] = est.stats_["streamed_bytes"] / 1e9
                d["gbs_on_bytes_moved"] = est.stats_["streamed_bytes"] / (ms * 1e-3) / 1e9
            d["R2Y_last"] = float(est.R2Y[-1])
            if alg == "stream":
                # per-class CUDA-event profile of one more fit (events between kernels cost a little: not the timed fit)
                est.fit(arg_of(cls, Xs), Y, profile=True)
                d["profile_ms"] = {k: round(v["ms"], 4) for k, v in est.profile_.items() if v["launches"]}
                d["profile_launches"] = {k: int(v["launches"]) for k, v in est.profile_.items() if v["launches"]}
            line[alg] = d
            del est
        if c.get("cv"):
            arg = Xs if len(Xs) > 1 else Xs[0]
            q2y_sweep(arg, Y, 2, n_splits=2)                    # warm-up
            line["cv_sweep"] = {"folds": c["cv"], "note": "one R-component fit per fold (nested components), folds as 0/1 row weights"}
            for alg in ("stream", "covariance"):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                q2 = q2y_sweep(arg, Y, c["R"], n_splits=c["cv"], seed=0, algorithm=alg)
                torch.cuda.synchronize()
                line["cv_sweep"][alg] = {"s_wall": time.perf_counter() - t0, "q2y": [float(v) for v in q2]}
        if not args.no_cpu and c["cpu_rows"]:
            # the oracle port (numpy restatement of the reference's fit, incl. its dense R2 re-evaluation) on the
            # first cpu_rows rows of the same data
            from oracle import tpls_oracle as orc
            nr = c["cpu_rows"]
            Xh = [x[:nr].cpu().numpy() for x in Xs]
            Yh = Y[:nr].cpu().numpy()
            Rc = min(c["R"], 3)
            t0 = time.perf_counter()
            st = orc.fit([x.copy() for x in Xh], Yh.copy(), Rc, r2_mode="reference")
            dt = time.perf_counter() - t0
            trips = int(st["trips"].sum())
            b = sum(elem * nr * int(np.prod(d)) for d in c["dims"]) * (2.0 * trips + Rc + 2)
            line["cpu_oracle"] = {"rows": nr, "components": Rc, "seconds": dt, "trips": trips, "gbs": b / dt / 1e9,
                                  "cores": len(os.sched_getaffinity(0))}
        print(json.dumps(line), file=real_stdout, flush=True)
        del Xs, Y
        trim_memory()
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
