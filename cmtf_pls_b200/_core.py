"""Host-side plumbing shared by :class:`tPLS` and :class:`ctPLS`: input
validation with the reference's error behaviour, engine/communicator caching,
fetching the fitted state, and the model-side inputs of ``transform``.

All arithmetic on X happens in the CUDA library; what is done here in numpy is
bookkeeping on replicated, loading-sized arrays (kron of loading vectors, the
final ``scores @ coef_ @ Q.T`` of predict).
"""

from __future__ import annotations

import os
from functools import reduce

import numpy as np

from . import _engine

_engines: dict = {}


def _is_torch(a) -> bool:
    return type(a).__module__.split(".")[0] == "torch"


def _device_of(arrays, default=None) -> int:
    for a in arrays:
        if _is_torch(a) and a.is_cuda:
            return a.device.index if a.device.index is not None else 0
    if default is not None:
        return int(default)
    try:
        import torch
        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except Exception:
        pass
    return 0


def _torch_stream_handle(device: int) -> int:
    """cudaStream_t of torch's current stream on ``device`` so that our kernels are
    ordered after whatever produced the input tensors; the legacy default stream
    (handle 0) is passed as cudaStreamLegacy (0x1).  0 = no torch / no CUDA: the
    library then creates a private stream."""
    try:
        import torch
        if torch.cuda.is_available():
            hnd = int(torch.cuda.current_stream(device).cuda_stream)
            return hnd if hnd != 0 else 1
    except Exception:
        pass
    return 0


def get_engine(device: int) -> _engine.Engine:
    """ONE engine per device (it caches X-sized working buffers); it follows torch's current stream from call
    to call instead of being bound to the stream it was created under."""
    stream = _torch_stream_handle(device)
    eng = _engines.get(device)
    if eng is None or eng.h is None:
        eng = _engine.Engine(device, stream or None)
        _engines[device] = eng
    else:
        eng.set_stream(stream)
    return eng


def _ensure_comm(eng, group):
    """Create the NCCL communicator of ``group`` (a torch.distributed process
    group, or True for the default one) once per engine."""
    import torch.distributed as dist
    pg = None if group is True else group
    world = dist.get_world_size(pg)
    rank = dist.get_rank(pg)
    # readiness lives on the ENGINE (the communicator and the exchange buffers belong to its handle): a recreated
    # engine, another group or another world size sets the communicator up again
    key = (id(pg) if pg is not None else 0, rank, world)
    if eng.comm_key == key and eng.comm_world == world:
        return rank, world
    if world <= 1:
        eng.init_comm(b"\0" * 128, 0, 1)
        eng.comm_key, eng.comm_world, eng.exchange = key, 1, "none"
        return rank, world
    box = [eng.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(pg, 0) if pg is not None else 0, group=pg)
    eng.init_comm(box[0], rank, world)
    eng.exchange = "nccl"
    # one-shot peer-memory exchange for the per-trip all-reduces (NCCL stays the fallback)
    if os.environ.get("TPLS_NO_XCHG", "") == "":
        ok, why = True, ""
        try:
            mine = eng.xchg_handle()
        except Exception as exc:  # noqa: BLE001
            mine, ok, why = b"\0" * 64, False, str(exc)
        handles = [None] * world
        dist.all_gather_object(handles, mine, group=pg)
        if ok:
            try:
                eng.xchg_open(handles)
            except Exception as exc:  # noqa: BLE001 -- e.g. CUDA IPC unavailable in this container
                ok, why = False, str(exc)
        oks = [None] * world
        dist.all_gather_object(oks, ok, group=pg)
        if not all(oks):  # all ranks or none: a mixed set-up would deadlock
            eng.xchg_open(None)
            if rank == 0:
                import warnings
                warnings.warn(f"peer-memory exchange unavailable ({why or 'on another rank'}); NCCL is used for every all-reduce")
        else:
            eng.exchange = "peer"
    eng.comm_key, eng.comm_world = key, world
    return rank, world


def _ensure_single(eng):
    eng.init_comm(b"\0" * 128, 0, 1)
    eng.comm_key, eng.comm_world, eng.exchange = None, 1, "none"


def weak_ref(a):
    import weakref
    try:
        return weakref.ref(a)
    except TypeError:   # an object that cannot be weakly referenced (e.g. a list): no reference at all
        return None


def isnan_of(ref, what):
    """NaN positions of a weakly referenced training array, or a clear error when it is gone."""
    a = ref() if ref is not None else None
    if a is None:
        raise RuntimeError(
            f"{what}: the training array is no longer alive (the estimator references it weakly and does not "
            "pickle it); np.isnan of the data you fitted on gives the same mask")
    if _is_torch(a):
        return a.isnan().cpu().numpy()
    return np.isnan(np.asarray(a))


def as_input(a, what):
    """numpy array or torch tensor, C-contiguous; anything else goes through np.asarray."""
    if _is_torch(a):
        return a if a.is_contiguous() else a.contiguous()
    a = np.asarray(a)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    return np.ascontiguousarray(a)


def y_as_f64_2d(Y):
    """Y as float64 (n, m); a 1-D Y becomes (n, 1) (tpls.py:48-49)."""
    if _is_torch(Y):
        import torch
        Y = Y.to(torch.float64)
        if Y.ndim == 1:
            Y = Y.reshape(-1, 1)
        return Y.contiguous()
    Y = np.asarray(Y, dtype=np.float64)
    if Y.ndim == 1:
        Y = Y.reshape(-1, 1)
    return np.ascontiguousarray(Y)


def np_dtype_of(a):
    return np.dtype(str(a.dtype).replace("torch.", ""))


def run_fit(Xs, Y, n_components, tol, max_iter, device=None, group=None, overwrite=False, flags=0, profile=False,
            row_weights=None, algorithm="stream"):
    """Upload (or adopt) the shards, run the device fit, fetch the state.

    Returns a dict with T, W (list per tensor of loading matrices), U, Q, coef,
    R2X (list), R2Y, X_mean (list), Y_mean, has_miss (list), trips, stats.
    """
    import time
    t_host = [time.perf_counter()]
    Xs = [as_input(X, "X") for X in Xs]
    Y2 = y_as_f64_2d(Y)
    dev = _device_of(Xs + [Y2], device)
    eng = get_engine(dev)
    if group is not None and group is not False:
        _, world = _ensure_comm(eng, group)
        if eng.comm_world != world:
            raise _engine.TplsError(f"engine communicator spans {eng.comm_world} ranks, the process group {world}")
    elif eng.comm_world != 1:
        _ensure_single(eng)   # a sharded fit came before on this device: back to a single-GPU handle
    n, m = int(Y2.shape[0]), int(Y2.shape[1])
    R = int(n_components)
    for i, X in enumerate(Xs):
        fl = _engine.X_MAY_OVERWRITE if (overwrite and _is_torch(X) and X.is_cuda) else 0
        eng.set_x(i, X, fl)
    eng.set_y(Y2)
    if row_weights is not None:
        w = row_weights if _is_torch(row_weights) else np.ascontiguousarray(row_weights, dtype=np.float64)
        eng.set_row_weights(w)
    if algorithm == "covariance":
        flags |= _engine.FIT_COVARIANCE
    elif algorithm != "stream":
        raise ValueError("algorithm must be 'stream' or 'covariance'")
    if profile:
        flags |= _engine.FIT_PROFILE
    try:
        t_host.append(time.perf_counter())
        eng.fit(len(Xs), R, tol, max_iter, flags)
        t_host.append(time.perf_counter())
        out = dict(
            T=eng.x_factor(0, 0, n, R),
            W=[[eng.x_factor(i, k, int(X.shape[k]), R) for k in range(1, X.ndim)] for i, X in enumerate(Xs)],
            U=eng.y_factor(0, n, R),
            Q=eng.y_factor(1, m, R),
            coef=eng.coef(R),
            R2X=[eng.r2x(i, R) for i in range(len(Xs))],
            R2Y=eng.r2y(R),
            X_mean=[eng.x_mean(i, tuple(int(s) for s in X.shape[1:]), np_dtype_of(X)) for i, X in enumerate(Xs)],
            Y_mean=eng.y_mean(m),
            has_miss=[eng.has_missing(i) for i in range(len(Xs))],
            trips=eng.trips(R),
            converged=eng.converged(R),
            stats=eng.stats(),
            profile=LazyProfile(eng) if profile else None,
            device=dev,
        )
        # host-side wall clock of the three stages of this call (ms): staging the inputs, the device fit
        # (returns when the GPU is done), fetching the fitted state
        t_host.append(time.perf_counter())
        out["stats"]["exchange"] = eng.exchange
        out["stats"]["host_ms"] = dict(stage=1e3 * (t_host[1] - t_host[0]), fit=1e3 * (t_host[2] - t_host[1]),
                                       fetch=1e3 * (t_host[3] - t_host[2]))
    finally:
        eng.release_data()
    return out


class LazyProfile:
    """Per-class CUDA-event profile of the profiled fits since the last read (`tpls_get_profile`).  The event
    queries (a few microseconds for each of the ~2000 launches of a fit) are made when it is read, not inside
    ``fit``; reading it after several profiled fits on the same device gives their sums."""

    def __init__(self, eng):
        self._eng = eng
        self._value = None

    def get(self):
        if self._value is None:
            self._value = self._eng.profile()
            self._eng = None
        return self._value


def kron_rows(loadings, R):
    """(R, P) matrix whose row a is kron(w_2[:, a], w_3[:, a], ...)."""
    return np.ascontiguousarray(np.stack([reduce(np.kron, [w[:, a] for w in loadings]) for a in range(R)]))


def run_transform(Xs, means, loadings_per_tensor, R, device=None):
    """Scores (n_new, R) of new data through the stored loadings."""
    Xs = [as_input(X, "X") for X in Xs]
    dev = _device_of(Xs, device)
    eng = get_engine(dev)
    wk = [kron_rows(ws, R) for ws in loadings_per_tensor]
    mm = [np.ascontiguousarray(np.asarray(mu, dtype=np_dtype_of(X)).reshape(-1)) for mu, X in zip(means, Xs)]
    # constants of the read-only path (complete data): the deflation recurrence on the scores alone needs
    # <mean_l, kron_l[a]> and <kron_l[b], kron_l[a]>, averaged over the coupled tensors like the scores
    c = np.mean([w @ m.astype(np.float64) for w, m in zip(wk, mm)], axis=0)
    G = np.mean([w @ w.T for w in wk], axis=0)
    if not (np.all(np.isfinite(c)) and np.all(np.isfinite(G))):  # e.g. a training column that was all NaN
        c = G = None
    else:
        c, G = np.ascontiguousarray(c), np.ascontiguousarray(G)
    return eng.transform(Xs, mm, wk, c, G)


def y_scores(Y, Y_mean, Y_shape, X_scores, coef, Q):
    """The Y branch of transform (tpls.py:167-184): small, replicated, host-side."""
    if _is_torch(Y):
        Y = Y.detach().cpu().numpy()
    Y = np.array(Y, dtype=np.float64, copy=True)
    if (Y.ndim != 1) and (Y.ndim != 2):
        raise ValueError("Only a matrix (2-mode tensor) Y is allowed.")
    if Y.ndim == 1:
        Y = Y.reshape((-1, 1))
    if tuple(Y_shape[1:]) != tuple(Y.shape[1:]):
        raise ValueError(f"Training Y has shape {tuple(Y_shape)}, while the new Y has shape {Y.shape}")
    Y -= Y_mean
    R = X_scores.shape[1]
    out = np.zeros((Y.shape[0], R))
    for a in range(R):
        out[:, a] = Y @ Q[:, a]
        Y -= X_scores @ coef[:, [a]] @ Q[:, [a]].T
    return out


def run_reconstruct(factors, mean, device=None):
    """sum_r a_r o b_r o ... + mean (util.py:18-20, tpls.py:188-189) on the device: one rank-R outer-product
    writer over the (N, P) result; only the (R, P) Kronecker rows are formed on the host."""
    R = factors[0].shape[1]
    eng = get_engine(_device_of([], device))
    wk = kron_rows(factors[1:], R)
    mu = None if mean is None else np.ascontiguousarray(np.asarray(mean).reshape(-1))
    out = eng.reconstruct(factors[0], wk, mu)
    return out.reshape([f.shape[0] for f in factors])
