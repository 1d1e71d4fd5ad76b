"""TEST INFRASTRUCTURE ONLY -- restated tensorly.cp_tensor leaves (see
../__init__.py).  Call sites in the reference: cmtf_pls/synthetic.py:26-29,67-71
(CPTensor, cp_to_tensor); tests/test_tpls.py:43-44 (cp_normalize)."""

import numpy as np


class CPTensor:
    """(weights, factors) pair behaving like a 2-sequence, with ``.rank`` and
    ``.shape``; arbitrary attributes may be attached (the reference hangs
    ``y_factor`` on it, synthetic.py:27,68)."""

    def __init__(self, cp_tensor):
        weights, factors = cp_tensor
        factors = list(factors)
        self.rank = int(np.shape(factors[0])[1])
        self.shape = tuple(int(np.shape(f)[0]) for f in factors)
        if weights is None:
            weights = np.ones(self.rank, dtype=np.asarray(factors[0]).dtype)
        self.weights = weights
        self.factors = factors

    def __getitem__(self, index):
        if index == 0:
            return self.weights
        if index == 1:
            return self.factors
        raise IndexError(index)

    def __iter__(self):
        yield self.weights
        yield self.factors

    def __len__(self):
        return 2


def _unpack(cp):
    weights, factors = cp
    if weights is None:
        weights = np.ones(np.shape(factors[0])[1], dtype=np.asarray(factors[0]).dtype)
    return weights, list(factors)


def cp_to_tensor(cp, mask=None):
    """Dense tensor  sum_r w_r * a_r o b_r o c_r ...  (mode 0 slowest, C
    order).  Fixed by mathematics."""
    weights, factors = _unpack(cp)
    shape = tuple(f.shape[0] for f in factors)
    lead = factors[0] * weights  # (I0, R)
    rest = np.ones((1, lead.shape[1]), dtype=lead.dtype)
    for f in factors[1:]:
        rest = (rest[:, None, :] * f[None, :, :]).reshape(-1, lead.shape[1])
    return (lead @ rest.T).reshape(shape)


def cp_normalize(cp):
    """Push the column norms of every factor into the weights.

    tensorly 0.9.0 semantics, restated: the incoming weights are first folded
    into factor 0; each factor's columns are divided by their l2 norm (a zero
    norm is left as a zero column, dividing by 1 instead); the weights become
    the product of those norms."""
    weights, factors = _unpack(cp)
    rank = factors[0].shape[1]
    out = []
    new_w = np.ones(rank, dtype=np.result_type(factors[0].dtype, np.float64))
    for i, f in enumerate(factors):
        if i == 0:
            f = f * weights
        scales = np.sqrt(np.sum(np.abs(f) ** 2, axis=0))
        safe = np.where(scales == 0, np.ones_like(scales), scales)
        new_w = new_w * scales
        out.append(f / safe[None, :])
    return CPTensor((new_w, out))


def cp_norm(cp):
    """Frobenius norm of the CP tensor from the factor Gram matrices."""
    weights, factors = _unpack(cp)
    gram = np.ones((len(weights), len(weights)))
    for f in factors:
        gram = gram * (f.T @ f)
    gram = gram * np.outer(weights, weights)
    return np.sqrt(np.abs(np.sum(gram)))
