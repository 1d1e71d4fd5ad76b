"""Benchmark of the tensor-PLS fit path (BASELINE.json metric: tPLS fit time and
effective X-stream GB/s against the HBM roofline, 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is ONE complete fit (all components) of the workload on synthetic data
already resident in HBM.  Workload = BASELINE.json configs[3]: a coupled pair of
1M x 64 x 64 fp32 tensors sharing the sample mode, Y 1M x 4, 10 components; the
sample mode is sharded over the N ranks (strong scaling: the total stays 1M).
M, the latent rank and the noise level are not fixed by BASELINE.json; this
file uses M=4, L=12, error=1.0 and prints them in `config`.

`value`   = effective X-stream bandwidth  B_alg / t_fit  summed over ranks, with
            B_alg = sum_l s*N*P_l*(2*sum(trips) + R + 2)  (SURVEY.md §8d);
`e2e`     = the same metric through the estimator API with HOST (pinned) arrays:
            host->device copy of X and Y and device->host copy of the fitted state
            are inside the timed region;
`roofline`= the dominant streaming kernel's algorithmic bytes per launch over its
            mean launch duration (CUDA events on the launching stream, collected
            during the timed steps) against MEASURED_PEAKS.json;
`cpu_baseline` = the oracle port of the reference's numpy fit on a bounded
            sample of the same workload, on the box's host cores.
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "tPLS fit effective X-stream bandwidth (B_alg / fit time)"
UNIT = "GB/s"
N_TOTAL = 1_000_000
DIMS = (64, 64)
M, LATENT, ERROR, R = 4, 12, 1.0, 10
CPU_SAMPLE_ROWS = int(os.environ.get("CMTF_BENCH_CPU_ROWS", "4096"))   # rows of the bounded CPU sample (tests shrink it)
CPU_SAMPLE_R = 3
# golden cases of the parity gate (fp64; tests/golden/*.npz): coupled pair, NaNs, 4-way X, 30 responses (the path with
# an explicit ||u_old - u_new||^2 exchange), a non-converging 4-way case
PARITY_CASES = ["ct_90x32x16_90x24_m4_r5", "t3_miss_70x12x8_m4_r4", "t4_60x8x6x4_m3_r4", "t2_same_xy_40x30_r4",
                "t4_maxiter_20x6x5x4_m5_r2"]
PARITY_TOL = 1e-8


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------
GEN_BLOCK = 15625  # rows per generator block: 1M / 64, so 1/2/4/8-way shards start on block boundaries


def make_shard_device(row_start, n_rows, device):
    """CP-structured coupled pair in the spirit of import_synthetic (synthetic.py:59-77):
    shared scores T, per-tensor mode factors, Gaussian noise; generated on the device
    (a 32.8 GB numpy draw is impractical).  Mode factors come from seed 215; the rows are
    drawn block by block from seeds that depend only on the GLOBAL block index, so the
    data set is the same whatever the number of ranks."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(215)
    yf = torch.randn(M, LATENT, generator=g, device=device, dtype=torch.float64)
    modes = [[torch.randn(d, LATENT, generator=g, device=device, dtype=torch.float64) for d in DIMS] for _ in range(2)]
    krs = [(a[:, None, :] * b[None, :, :]).reshape(-1, LATENT).float() for a, b in modes]      # (P, L) each
    Xs = [torch.empty(n_rows, *DIMS, dtype=torch.float32, device=device) for _ in modes]
    Y = torch.empty(n_rows, M, dtype=torch.float64, device=device)
    row_stop = row_start + n_rows
    for blk in range(row_start // GEN_BLOCK, (row_stop + GEN_BLOCK - 1) // GEN_BLOCK):
        g.manual_seed(1000 + blk)
        b0 = blk * GEN_BLOCK
        lo, hi = max(b0, row_start), min(b0 + GEN_BLOCK, row_stop)
        T = torch.randn(GEN_BLOCK, LATENT, generator=g, device=device, dtype=torch.float64)
        Yb = T @ yf.T + ERROR * torch.randn(GEN_BLOCK, M, generator=g, device=device, dtype=torch.float64)
        Y[lo - row_start:hi - row_start] = Yb[lo - b0:hi - b0]
        for X, kr in zip(Xs, krs):
            blkx = T.float() @ kr.T
            blkx += ERROR * torch.randn(GEN_BLOCK, kr.shape[0], generator=g, device=device, dtype=torch.float32)
            X.view(n_rows, -1)[lo - row_start:hi - row_start] = blkx[lo - b0:hi - b0]
    return Xs, Y


def make_sample_host(n_rows, seed=215):
    rng = np.random.default_rng(seed)
    T = rng.normal(size=(n_rows, LATENT))
    yf = rng.normal(size=(M, LATENT))
    Xs = []
    for _ in range(2):
        a, b = rng.normal(size=(DIMS[0], LATENT)), rng.normal(size=(DIMS[1], LATENT))
        X = np.einsum("ir,jr,kr->ijk", T, a, b) + rng.normal(0, ERROR, size=(n_rows,) + DIMS)
        Xs.append(X.astype(np.float32))
    Y = T @ yf.T + rng.normal(0, ERROR, size=(n_rows, M))
    return Xs, Y


def alg_bytes(n_rows, trips_total, n_comp):
    p = DIMS[0] * DIMS[1]
    return 2 * 4.0 * n_rows * p * (2.0 * trips_total + n_comp + 2)


# --------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for row in self.rows:
            f = [c.strip() for c in row.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = [s for s in sm if s >= 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(local):
    """Pin this rank's threads to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host memory is
    allocated (first touch decides where the pages live): with 8 ranks on a two-socket host a rank whose staging
    buffers sit on the far socket uploads at ~25 GB/s instead of ~55 GB/s.  Best effort; returns what it did."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return {"numa_node": None, "why": "the platform reports no NUMA affinity for the GPU"}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"numa_node": node, "why": "none of the node's CPUs is in this process's affinity mask"}
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus), "pci": bus}
    except Exception as exc:  # noqa: BLE001
        return {"numa_node": None, "why": repr(exc)[:120]}


# --------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's numpy fit on a bounded sample
# --------------------------------------------------------------------------
def cpu_fit_once(Xs, Y):
    from oracle import tpls_oracle as orc
    t0 = time.perf_counter()
    st = orc.fit([x.copy() for x in Xs], Y.copy(), CPU_SAMPLE_R, r2_mode="reference")
    dt = time.perf_counter() - t0
    trips = int(st["trips"].sum())
    return dt, trips, alg_bytes(Xs[0].shape[0], trips, CPU_SAMPLE_R)


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        n = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        n = 1
    return int(n), len(os.sched_getaffinity(0))


def cpu_extrapolation(seconds, rate_gbs, trips_full=None):
    """a full-size CPU figure can only be a linear-in-N EXTRAPOLATION of the bounded sample and must
    be labelled as such (the reference would need hours and > 100 GB of host RAM for 1M rows)."""
    out = {"label": "EXTRAPOLATION, linear in N, from the bounded sample -- not a measurement",
           "same_fit_at_1M_rows_s": seconds * N_TOTAL / CPU_SAMPLE_ROWS,
           "basis": f"{CPU_SAMPLE_R}-component fit of {CPU_SAMPLE_ROWS} rows took {seconds:.2f} s"}
    if trips_full is not None and rate_gbs > 0:
        out["full_workload_s"] = alg_bytes(N_TOTAL, trips_full, R) / (rate_gbs * 1e9)
        out["full_workload"] = f"{R} components, {trips_full} inner trips (the GPU fit's count), at the sample's B_alg rate"
    return out


def cpu_baseline_block(value, cores, affinity):
    return {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": (f"oracle/tpls_oracle.fit (numpy restatement of cmtf_pls ctPLS.fit incl. its dense R2X/R2Y "
                       f"re-evaluation) on 2 x [{CPU_SAMPLE_ROWS} x 64 x 64] fp32 + Y {CPU_SAMPLE_ROWS} x {M}, "
                       f"{CPU_SAMPLE_R} components; BLAS threads {cores}, affinity {affinity}; the contraction is "
                       f"single-threaded numpy einsum as in the reference")}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    Xs, Y = make_sample_host(CPU_SAMPLE_ROWS)
    cores, aff = cpu_threads()
    for _ in range(args.warmup):
        cpu_fit_once(Xs, Y)
    t_tot, b_tot, trips = 0.0, 0.0, 0
    for _ in range(args.steps):
        dt, tr, b = cpu_fit_once(Xs, Y)
        t_tot += dt
        b_tot += b
        trips = tr
    val = b_tot / t_tot / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[3] shape (coupled pair x64x64 fp32, M=4), bounded CPU sample",
                   "rows": CPU_SAMPLE_ROWS, "components": CPU_SAMPLE_R, "latent": LATENT, "error": ERROR, "trips": trips},
        "cpu_baseline": dict(cpu_baseline_block(val, cores, aff), extrapolation=cpu_extrapolation(t_tot / args.steps, val)),
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from cmtf_pls_b200 import ctPLS

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else {"numa_node": None, "why": "single rank: not bound"}
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = True
    n_total = args.rows or N_TOTAL
    from cmtf_pls_b200.sharding import row_block
    row_lo, row_hi = row_block(n_total, rank, world)
    n_loc = row_hi - row_lo

    Xs, Y = make_shard_device(row_lo, n_loc, dev)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity gate (before anything is timed): golden cases minted from the reference's unmodified source,
    #      fitted with their rows sharded over the N ranks, against the stored single-process result.  The files hold
    #      the answers, so no oracle code runs here.  A failure ends the run with a non-zero exit code.
    parity = None
    if not args.no_parity:
        from cmtf_pls_b200.selfcheck import sharded_golden_check

        def gather_rows(local_rows, n_all, lo, hi):
            full = torch.zeros((n_all,) + tuple(local_rows.shape[1:]), dtype=torch.float64, device=dev)
            full[lo:hi] = torch.from_numpy(np.ascontiguousarray(local_rows)).to(dev)
            dist.all_reduce(full)
            return full.cpu().numpy()

        parity = sharded_golden_check(PARITY_CASES, local, process_group=group, rank=rank, world=world, gather=gather_rows)
        parity["tolerance"] = PARITY_TOL
        parity["ok"] = bool(parity["trips_equal"] and parity["max_err"] < PARITY_TOL)
        if not parity["ok"]:
            if rank == 0:
                print(json.dumps({"parity_check": parity, "error": "sharded fit does not match the golden vectors"}), flush=True)
            if world > 1:
                dist.destroy_process_group()
            sys.exit(3)

    # ---- the two small BASELINE configurations (one GPU): their inner trips run in the resident trip loop (DESIGN.md
    #      §3b), one launch per component; device time of a fit (CUDA events inside the library), best of five.  They
    #      are latency-bound, i.e. proportional to the SM clock: measured here, before the streaming workload has
    #      pulled the GPU into its power cap (after it they take ~10-20 % longer at ~1.43 GHz) ----
    small = None
    if world == 1 and not args.quick:
        try:
            small = small_configs(local)
        except Exception as exc:  # noqa: BLE001
            small = {"error": str(exc)[:200]}

    est = ctPLS(R, device=local, process_group=group)
    for _ in range(args.warmup):
        est.fit(Xs, Y)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    barrier()
    ev0.record()
    for _ in range(args.steps):
        est.fit(Xs, Y)                    # one CUDA-graph launch per fit: the host takes no part in the inner loops
        launches += est.stats_["kernel_launches"]
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    stats_timed = dict(est.stats_)
    # how long every rank waited inside the exchange kernels of the last timed fit: the rank that waits least sets the pace
    xw = torch.tensor([stats_timed.get("xchg_wait_ms", 0.0), stats_timed.get("xchg_ms", 0.0), stats_timed.get("fit_ms", 0.0)],
                      dtype=torch.float64, device=dev)
    xw_all = [torch.zeros_like(xw) for _ in range(world)]
    if world > 1:
        dist.all_gather(xw_all, xw)
    else:
        xw_all = [xw]
    xchg_diag = {"exchanges_per_fit": int(stats_timed.get("xchg_count", 0)),
                 "wait_ms_per_rank": [round(float(t[0]), 2) for t in xw_all],
                 "start_to_synchronised_ms_per_rank": [round(float(t[1]), 2) for t in xw_all],
                 "fit_ms_per_rank": [round(float(t[2]), 2) for t in xw_all],
                 "note": "summed over the exchanges of one fit, measured by CTA 0 of each exchange kernel with %globaltimer"}
    # per-kernel-class durations: the SAME fit enqueued kernel by kernel by the host (a graph cannot carry an event
    # pair per launch), CUDA events around every launch on the launching stream, K more steps right after the timed ones
    evp0, evp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evp0.record()
    for _ in range(args.steps):
        est.fit(Xs, Y, profile=True)
    evp1.record()
    barrier()
    ms_profiled = evp0.elapsed_time(evp1)
    prof = est.profile_                    # the event pairs of all K fits are queried here
    clocks = sampler.stop() if rank == 0 else None
    trips = int(est.n_iter_.sum())
    tt = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    bb = torch.tensor([alg_bytes(n_loc, trips, R)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(bb, op=dist.ReduceOp.SUM)
    ms_step = tt.item() / args.steps
    value = bb.item() / (ms_step * 1e-3) / 1e9
    fit_ms_device = est.stats_["fit_ms"]
    host_ms_last = est.stats_.get("host_ms")

    # ---- the same fit in covariance mode, reported beside the headline, never as it:
    #      it moves ~2.5 passes of X per component instead of two per inner trip, so its rate is quoted on the
    #      bytes it ACTUALLY streamed ----
    cov = None
    try:
        if args.quick:
            raise RuntimeError("skipped (--quick)")
        est_c = ctPLS(R, device=local, process_group=group, algorithm="covariance")
        for _ in range(2):     # two warm-ups: the second fit of an estimator still pins fresh result buffers
            est_c.fit(Xs, Y)
        barrier()
        evc0, evc1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        evc0.record()
        for _ in range(args.steps):
            est_c.fit(Xs, Y)
        evc1.record()
        barrier()
        tc = torch.tensor([evc0.elapsed_time(evc1) / args.steps], dtype=torch.float64, device=dev)
        bc = torch.tensor([est_c.stats_["streamed_bytes"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
            dist.all_reduce(bc, op=dist.ReduceOp.SUM)
        cov = {"ms_per_step": tc.item(), "bytes_streamed_per_step": bc.item(),
               "gbs_on_bytes_actually_moved": bc.item() / (tc.item() * 1e-3) / 1e9,
               "trips": est_c.n_iter_.tolist(), "ran_covariance_loop": bool(est_c.stats_["covariance_mode"]),
               "max_abs_diff_R2Y_vs_streaming": float(np.max(np.abs(est_c.R2Y - est.R2Y))),
               "speedup_vs_streaming_fit": ms_step / tc.item()}
        del est_c
    except Exception as exc:  # noqa: BLE001
        cov = {"error": repr(exc)[:200]}

    # ---- transform of the resident shard through the fitted model: complete data is read in
    #      place ONCE (all R projections in one pass on the fp64 tensor-core path + the score recurrence); the
    #      (n, R) scores come back to the host.  Beside it: the masked (sequential) path on a NaN-carrying copy of
    #      the first rows, and the device reconstruction X_hat = T W + mean of a block of rows ----
    xform = None
    try:
        if args.quick:
            raise RuntimeError("skipped (--quick)")
        from cmtf_pls_b200 import _core as _c
        est.transform(Xs)
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            est.transform(Xs)
        barrier()
        dt = (time.perf_counter() - t0) / 3
        st_x = _c.get_engine(local).stats()
        tt2 = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt2, op=dist.ReduceOp.MAX)
        passes = 1 if st_x["last_transform_path"] == 1 else 2 * R + 1
        moved = 2 * 4.0 * n_total * 4096 * passes
        xform = {"s_per_call": tt2.item(), "rows_per_s": n_total / tt2.item(), "gbs_on_bytes_moved": moved / tt2.item() / 1e9,
                 "passes": passes, "kernel_launches": int(st_x["kernel_launches"]),
                 "fp64_fma_per_call": 2.0 * n_total * 4096 * R,
                 "includes": "D2H of the scores, host clock; R=10 makes the pass as heavy in fp64 FMAs (P*R per row, padded "
                             "to 16 components on the DMMA path) as in HBM bytes"}
        nm = min(n_loc, 100_000)
        Xm = [x[:nm].clone() for x in Xs]
        for x in Xm:
            x.view(-1)[::17] = float("nan")
        est.transform(Xm)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        est.transform(Xm)
        torch.cuda.synchronize()
        dtm = time.perf_counter() - t0
        xform["masked"] = {"rows": nm, "nan_fraction": 1 / 17, "s_per_call": dtm, "rows_per_s": nm / dtm,
                           "gbs_on_bytes_moved": 2 * 4.0 * nm * 4096 * (3 * R + 1) / dtm / 1e9,
                           "passes": "centre r+w, then per component a masked projection and a deflation r+w"}
        del Xm
        nr = min(n_loc, 20_000)
        t0 = time.perf_counter()
        xr = _c.run_reconstruct([est.factor_T[:nr]] + est.Xs_factors[0][1:], est.Xs_mean[0], device=local)
        dtr = time.perf_counter() - t0
        xform["reconstruct"] = {"rows": nr, "s_per_call": dtr, "out_gb": xr.nbytes / 1e9,
                                "note": "device writer + D2H of the (rows, 64, 64) float64 result, host clock"}
        del xr
    except Exception as exc:  # noqa: BLE001
        xform = {"error": repr(exc)[:300]}

    # ---- end to end through the estimator API with pinned host arrays ----
    e2e = None
    n_iter_resident = est.n_iter_.tolist()
    try:
        if args.e2e_steps <= 0 or args.quick:
            raise RuntimeError("skipped (--e2e-steps 0 / --quick)")
        Xh = [torch.empty(x.shape, dtype=x.dtype, pin_memory=True) for x in Xs]
        for h, d in zip(Xh, Xs):
            h.copy_(d)
        Yh = torch.empty(Y.shape, dtype=Y.dtype, pin_memory=True)
        Yh.copy_(Y)
        torch.cuda.synchronize()
        n_iter_resident = est.n_iter_.tolist()
        del Xs, est                                # free the resident shard before the host-array fits
        torch.cuda.empty_cache()
        Xn, Yn = [h.numpy() for h in Xh], Yh.numpy()
        est2 = ctPLS(R, device=local, process_group=group)
        est2.fit(Xn, Yn)                         # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            est2.fit(Xn, Yn)
        barrier()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        trips2 = int(est2.n_iter_.sum())
        t2 = torch.tensor([dt], dtype=torch.float64, device=dev)
        b2 = torch.tensor([alg_bytes(n_loc, trips2, R)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
            dist.all_reduce(b2, op=dist.ReduceOp.SUM)
        h2d = sum(x.nbytes for x in Xn) + Yn.nbytes
        d2h = 8 * (2 * n_loc * R + sum(DIMS) * 2 * R + M * R + R * R + 3 * R + M) + 4 * 2 * DIMS[0] * DIMS[1]
        e2e = {"value": b2.item() / t2.item() / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "s_per_step": t2.item(), "steps": args.e2e_steps,
               "numa": numa, "fit_ms_device_rank0": est2.stats_["fit_ms"],
               "h2d_gbs_rank0": (h2d / 1e9) / max(1e-9, t2.item() - est2.stats_["fit_ms"] * 1e-3 - 0.004),
               "timing": "host clock around est.fit(numpy pinned), barrier + synchronize on both sides, max over ranks"}
    except Exception as exc:  # noqa: BLE001
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": repr(exc)[:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_gbs()
    dom = max(("contract", "project", "deflate_contract"), key=lambda k: prof[k]["ms"])
    d = prof[dom]
    ach = d["bytes"] / d["launches"] / (d["ms"] / d["launches"] * 1e-3) / 1e9 if d["launches"] else None
    kernel_ms = sum(v["ms"] for v in prof.values())
    traffic = None
    try:  # DRAM bytes per launch from the committed ncu capture, scaled to this launch's algorithmic bytes
        with open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")) as f:
            tj = json.load(f)
        kk = tj["rowpass_kernel" if dom == "project" else "colpass_kernel_contract"]
        traffic = (kk["dram_read_bytes"] + kk["dram_write_bytes"]) / tj["algorithmic_bytes_per_launch"] * (d["bytes"] / max(1, d["launches"]))
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": {"contract": "colpass_kernel<PF_CONTRACT>", "project": "rowpass_kernel",
                                   "deflate_contract": "colpass_kernel<PF_DEFLATE|PF_WRITE|PF_CONTRACT|PF_SUMSQ>"}[dom],
        "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak if ach else None, "traffic": traffic,
        "peak_source": peak_src,
        "note": "read-only stream measured against a COPY (read+write) peak, so a fraction slightly above 1 is expected; "
                "traffic = ncu dram bytes per launch (profiles/r02_ncu_traffic.json) scaled to this launch size",
        "bytes_per_launch": d["bytes"] / max(1, d["launches"]), "ms_per_launch": d["ms"] / max(1, d["launches"]),
        "share_of_step": d["ms"] / (ms_profiled if world == 1 else kernel_ms),
        "measured": "CUDA events around every launch of %d profiled (host-enqueued) fits run right after the timed "
                    "(graph-launched) ones; profiled fit %.1f ms vs timed fit %.1f ms" % (args.steps, ms_profiled / args.steps, ms_total / args.steps),
        "per_class": {k: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps,
                          "gbs": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] > 0 and v["bytes"] > 0 else None}
                      for k, v in prof.items()},
    }
    cpu = None
    if world == 1 and not args.no_cpu:
        Xc, Yc = make_sample_host(CPU_SAMPLE_ROWS)
        cores, aff = cpu_threads()
        dt, tr, b = cpu_fit_once(Xc, Yc)
        cpu = cpu_baseline_block(b / dt / 1e9, cores, aff)
        cpu["seconds"] = dt
        cpu["trips"] = tr
        cpu["extrapolation"] = cpu_extrapolation(dt, b / dt / 1e9, trips)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "BASELINE configs[3]: coupled pair 2 x [1M x 64 x 64] fp32 + Y 1M x 4, 10 components, "
                               "sample mode sharded over the ranks",
                   "rows_total": n_total, "rows_per_gpu": n_loc, "responses": M, "latent": LATENT, "error": ERROR,
                   "components": R, "storage": "fp32 X, fp64 accumulators/vectors", "tol": 1e-8, "max_iter": 100,
                   "trips": n_iter_resident, "trips_total": trips,
                   "l2": "every pass streams 2 x %.1f GB per GPU, far larger than the 126 MB L2 (no flush needed)"
                         % (4.0 * n_loc * 4096 / 1e9),
                   "fraction_of_hbm_peak": value / (world * peak), "fit_ms_device_last": fit_ms_device,
                   "loop": "device-resident (one CUDA graph per fit, a WHILE node per component)" if stats_timed.get("graph_launches") else "host-enqueued trips",
                   "launches_per_trip": stats_timed.get("launches_per_trip"), "exchange": stats_timed.get("exchange"),
                   "launches_per_fit": stats_timed.get("kernel_launches"), "exchange_wait": xchg_diag if world > 1 else None,
                   "host_ms_last_fit": host_ms_last,
                   "covariance_mode": cov, "transform": xform, "small_configs": small},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        "parity_check": parity,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def small_configs(device):
    """BASELINE configs[0] (20 x 8 x 6, Y 20 x 1, 3 components) and configs[1] (10k x 32 x 16 coupled with 10k x 24, Y 10k x 4,
    5 components, fp64) on one GPU: synthetic CP-structured data drawn on the device, data resident in HBM."""
    import torch
    from cmtf_pls_b200 import ctPLS

    def draw(n, dims_list, m, latent, error, seed):
        g = torch.Generator(device="cuda")
        g.manual_seed(seed)
        T = torch.randn(n, latent, generator=g, device="cuda", dtype=torch.float64)
        yf = torch.randn(m, latent, generator=g, device="cuda", dtype=torch.float64)
        xs = []
        for dims in dims_list:
            kr = torch.ones(1, latent, device="cuda", dtype=torch.float64)
            for d in dims:
                f = torch.randn(d, latent, generator=g, device="cuda", dtype=torch.float64)
                kr = (kr[:, None, :] * f[None, :, :]).reshape(-1, latent)
            x = T @ kr.T + error * torch.randn(n, kr.shape[0], generator=g, device="cuda", dtype=torch.float64)
            xs.append(x.reshape(n, *dims).contiguous())
        y = T @ yf.T + error * torch.randn(n, m, generator=g, device="cuda", dtype=torch.float64)
        return xs, y

    out = {}
    for name, n, dims_list, m, latent, error, r in (("configs[0]", 20, [(8, 6)], 1, 3, 0.1, 3),
                                                    ("configs[1]", 10_000, [(32, 16), (24,)], 4, 8, 0.5, 5)):
        xs, y = draw(n, dims_list, m, latent, error, 7)
        est = ctPLS(r, device=device)
        est.fit(xs, y)
        ms = []
        for _ in range(5):
            est.fit(xs, y)
            ms.append(est.stats_["fit_ms"])
        trips = int(est.n_iter_.sum())
        out[name] = {"fit_ms": min(ms), "trips": trips, "us_per_trip": min(ms) / trips * 1e3,
                     "kernel_launches": int(est.stats_["kernel_launches"]), "resident_loops": int(est.stats_["resident_loops"]),
                     "components": r}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=0, help="override the total sample count (debugging)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the golden-vector parity gate (debugging)")
    ap.add_argument("--quick", action="store_true", help="headline fit only: no covariance / transform / e2e legs (A/B runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
