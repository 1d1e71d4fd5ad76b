"""TEST INFRASTRUCTURE ONLY -- restated tensorly.tenalg leaves (see
../__init__.py).  All three are fixed by mathematics."""

import numpy as np


def multi_mode_dot(tensor, matrix_or_vec_list, modes=None, skip=None, transpose=False):
    """Successive mode-n products, lowest mode first.  The reference only ever
    passes 1-D vectors (cmtf_pls/tpls.py:97-99,139-141,162-164;
    cmtf.py:107-111,157-161,194-198); contracting with a vector removes that
    mode, so later mode numbers shift down by one."""
    if modes is None:
        modes = range(len(matrix_or_vec_list))
    out = np.asarray(tensor)
    removed = 0
    for i, (op, mode) in enumerate(sorted(zip(matrix_or_vec_list, modes), key=lambda p: p[1])):
        if skip is not None and i == skip:
            continue
        op = np.asarray(op)
        ax = mode - removed
        if op.ndim == 1:
            out = np.tensordot(out, op, axes=([ax], [0]))
            removed += 1
        else:
            mat = op.T if transpose else op
            out = np.moveaxis(np.tensordot(out, mat, axes=([ax], [1])), -1, ax)
    return out


def outer(tensors):
    """N-way outer product of a list of vectors (cmtf_pls/tpls.py:109,142,165)."""
    out = np.asarray(tensors[0])
    for t in tensors[1:]:
        t = np.asarray(t)
        out = out.reshape(out.shape + (1,) * t.ndim) * t
    return out


def khatri_rao(matrices, weights=None, skip_matrix=None, reverse=False, mask=None):
    """Column-wise Kronecker product; the first kept matrix varies slowest
    (cmtf_pls/util.py:19 with skip_matrix=0)."""
    mats = [m for i, m in enumerate(matrices) if i != skip_matrix]
    if reverse:
        mats = mats[::-1]
    rank = mats[0].shape[1]
    out = np.ones((1, rank), dtype=np.result_type(*[m.dtype for m in mats]))
    for m in mats:
        out = (out[:, None, :] * m[None, :, :]).reshape(-1, rank)
    if weights is not None:
        out = out * np.reshape(weights, (1, -1))
    return out
