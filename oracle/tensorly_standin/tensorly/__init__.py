"""TEST INFRASTRUCTURE ONLY -- a numpy restatement of the handful of tensorly
0.9.0 leaf functions that meyer-lab/cmtf-pls calls.

Why this exists: the reference pins ``tensorly==0.9.0`` (requirements.lock:34,
pyproject.toml:10) and imports it at module top level (cmtf_pls/tpls.py:9-10,
cmtf.py:8-9, util.py:3-4, synthetic.py:2).  tensorly is not installed in this
image and there is no network, so the reference cannot be imported as shipped.
Putting this directory on ``sys.path`` lets ``/root/reference/cmtf_pls`` run
UNMODIFIED; only the leaves below are restated (from tensorly's published
behaviour -- its source is not vendored under /root/reference).

PARITY STATUS: every leaf except ``parafac`` is fixed by mathematics (see each
docstring).  ``parafac`` has algorithmic freedom when the tensor has >= 3 modes
(i.e. X has >= 4 modes): *parity unpinned* there; see decomposition/_cp.py.

Nothing in the product package (``cmtf_pls_b200``) may import this module.
"""

import numpy as np

from . import cp_tensor, tenalg  # noqa: F401  (reference does tl.cp_tensor.X)
from .cp_tensor import CPTensor, cp_to_tensor, cp_normalize  # noqa: F401

__version__ = "0.9.0+standin"


def fold(unfolded, mode, shape):
    """Inverse of the mode-``mode`` unfolding (row index = that mode, the other
    modes C-ordered along the columns).  Call site: cmtf_pls/util.py:20 (mode 0
    only, where it is a plain reshape)."""
    shape = tuple(shape)
    lead = (shape[mode],) + shape[:mode] + shape[mode + 1:]
    return np.moveaxis(np.reshape(unfolded, lead), 0, mode)


def unfold(tensor, mode):
    """Mode-``mode`` unfolding, tensorly convention (C order of the remaining
    modes)."""
    return np.reshape(np.moveaxis(tensor, mode, 0), (tensor.shape[mode], -1))


def norm(tensor, order=2, axis=None):
    """l2 (default) / l1 / inf norm, optionally along an axis.  Used by the
    reference's tests (tests/test_tpls.py:34,36) with ``axis=0``."""
    if order == "inf":
        return np.max(np.abs(tensor), axis=axis)
    if order == 1:
        return np.sum(np.abs(tensor), axis=axis)
    if order == 2:
        return np.sqrt(np.sum(np.abs(tensor) ** 2, axis=axis))
    return np.sum(np.abs(tensor) ** order, axis=axis) ** (1.0 / order)


def dot(a, b):
    """Matrix product.  Call sites: cmtf_pls/synthetic.py:30,73."""
    return np.dot(a, b)
