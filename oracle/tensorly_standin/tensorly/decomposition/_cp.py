"""TEST INFRASTRUCTURE ONLY -- restated ``tensorly.decomposition._cp.parafac``
for the one way the reference calls it (cmtf_pls/tpls.py:86, cmtf.py:100):

    parafac(Z, 1, tol=tol, init="svd", normalize_factors=True)[1]

tensorly 0.9.0's source is NOT under /root/reference and is not installed
here; this follows its published algorithm (ALS with an SVD/HOSVD start):

  init   for every mode, the leading ``rank`` left singular vectors of the
         mode unfolding (exact LAPACK SVD), each column sign-flipped so that
         its largest-|entry| is positive; the mode-0 block is scaled by the
         singular values; then ``cp_normalize`` (norms -> weights).
  sweep  for mode = 0..d-1:  factor <- MTTKRP(mode) / (w w^T * prod of the
         other factors' Gram matrices), the weights being part of both sides.
  error  after the last mode, sqrt|‖Z‖² + ‖Ẑ‖² − 2<Z,Ẑ>| / ‖Z‖ with <Z,Ẑ> taken
         from the last mode's MTTKRP and the freshly updated last factor.
  stop   from the second sweep on, when |err_prev − err| < tol; at most
         ``n_iter_max`` (100) sweeps.
  norm   ``cp_normalize`` at the END of a sweep that did not stop.

PARITY STATUS
  * Z a matrix (X has 3 modes): the SVD start is already the ALS fixed point,
    so the answer is the leading singular pair -- pinned by mathematics; any
    faithful reading of tensorly gives the same vectors to rounding.
  * Z with >= 3 modes (X has >= 4 modes): *parity unpinned*.  The stopping
    rule leaves the factors only ~sqrt(tol) converged, so the result depends on
    sweep-level details that could not be checked against the real package.
    The one detail we know to be uncertain is whether the sweep that triggers
    the stop is followed by a normalisation; ``NORMALIZE_ON_BREAK`` selects it
    (default False: break first, as recalled from 0.9.0).  The two readings
    differ by ~6e-8 element-relative on a 4-way fit and no
    reference test discriminates them.
"""

import numpy as np

from .. import unfold
from ..cp_tensor import CPTensor, cp_normalize, cp_norm
from ..tenalg import khatri_rao

NORMALIZE_ON_BREAK = False

# sweep counts of the most recent calls, for tests/diagnostics
last_sweeps = []


def _svd_flip_u(U):
    idx = np.argmax(np.abs(U), axis=0)
    signs = np.sign(U[idx, np.arange(U.shape[1])])
    return U * signs


def _leading_left_singular(matrix, k):
    full = k > min(matrix.shape)
    U, S, _ = np.linalg.svd(matrix, full_matrices=full)
    return _svd_flip_u(U[:, :k]), S[:k]


def initialize_cp(tensor, rank, normalize_factors):
    factors = []
    for mode in range(tensor.ndim):
        U, S = _leading_left_singular(unfold(tensor, mode), rank)
        if mode == 0:
            k = min(rank, S.shape[0])
            U = U.copy()
            U[:, :k] = U[:, :k] * S[:k]
        if U.shape[1] < rank:  # never reached for rank 1
            raise NotImplementedError("random completion of a short SVD basis")
        factors.append(U[:, :rank])
    cp = CPTensor((None, factors))
    if normalize_factors:
        cp = cp_normalize(cp)
    return cp


def parafac(tensor, rank, n_iter_max=100, init="svd", normalize_factors=False,
            tol=1e-8, **unused):
    if init != "svd":
        raise NotImplementedError("only init='svd' is restated")
    tensor = np.asarray(tensor, dtype=np.float64)
    weights, factors = initialize_cp(tensor, rank, normalize_factors)
    factors = list(factors)
    norm_tensor = np.sqrt(np.sum(tensor ** 2))
    rec_errors = []
    sweeps = 0
    for iteration in range(n_iter_max):
        sweeps += 1
        mttkrp = None
        for mode in range(tensor.ndim):
            gram = np.ones((rank, rank))
            for i, f in enumerate(factors):
                if i != mode:
                    gram = gram * (f.T @ f)
            gram = weights.reshape(-1, 1) * gram * weights.reshape(1, -1)
            kr = khatri_rao(factors, weights=weights, skip_matrix=mode)
            mttkrp = unfold(tensor, mode) @ kr
            factors[mode] = np.linalg.solve(gram.T, mttkrp.T).T
        if tol:
            fnorm = cp_norm((weights, factors))
            iprod = np.sum(np.sum(mttkrp * factors[-1], axis=0))
            err = np.sqrt(np.abs(norm_tensor ** 2 + fnorm ** 2 - 2 * iprod)) / norm_tensor
            rec_errors.append(err)
            if iteration >= 1 and abs(rec_errors[-2] - rec_errors[-1]) < tol:
                if NORMALIZE_ON_BREAK and normalize_factors:
                    weights, factors = cp_normalize((weights, factors))
                    factors = list(factors)
                break
        if normalize_factors:
            weights, factors = cp_normalize((weights, factors))
            factors = list(factors)
    last_sweeps.append(sweeps)
    del last_sweeps[:-64]
    return CPTensor((weights, factors))
