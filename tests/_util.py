"""Shared helpers for the parity tests: golden loading, per-component sign
alignment and column-wise Frobenius-relative error.  They live in
``cmtf_pls_b200.selfcheck`` (bench.py uses them too, for its sharded parity gate)."""

from cmtf_pls_b200.selfcheck import GOLDEN, aligned_errors, col_err, golden_cases, load_golden  # noqa: F401
