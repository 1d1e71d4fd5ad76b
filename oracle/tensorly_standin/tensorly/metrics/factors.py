"""TEST INFRASTRUCTURE ONLY -- restated ``congruence_coefficient`` (used only
by the reference's tests, tests/test_tpls.py:7,94-95)."""

import numpy as np
from scipy.optimize import linear_sum_assignment


def congruence_coefficient(matrix1, matrix2, absolute_value=True):
    """Best-permutation mean Tucker congruence between the columns of two
    factor matrices; returns (mean congruence, column permutation)."""
    a = matrix1 / np.linalg.norm(matrix1, axis=0)
    b = matrix2 / np.linalg.norm(matrix2, axis=0)
    c = a.T @ b
    if absolute_value:
        c = np.abs(c)
    rows, cols = linear_sum_assignment(-c)
    return c[rows, cols].mean(), list(cols)
