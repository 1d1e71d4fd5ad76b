"""Geometry sweep helper: times the three streaming operators on the c4 shard
shape under the TPLS_TILE_KB / TPLS_SMEM_KB / TPLS_CTAS_PER_SM overrides that
are set in the environment.  One line of output."""

import ctypes as C
import os
import sys

import torch

sys.path.insert(0, ".")
from cmtf_pls_b200._core import get_engine  # noqa: E402


def main():
    n = int(os.environ.get("TUNE_ROWS", 250_000))
    p = int(os.environ.get("TUNE_P", 4096))
    dt = torch.float32 if os.environ.get("TUNE_DTYPE", "f32") == "f32" else torch.float64
    masked = int(os.environ.get("TUNE_MASKED", 0))
    eng = get_engine(0)
    lib = eng.lib
    code = 0 if dt == torch.float32 else 1
    X = torch.randn(n, p, dtype=dt, device="cuda")
    if masked:
        X[torch.rand(n, p, device="cuda") < 0.2] = float("nan")
    u = torch.randn(n, dtype=torch.float64, device="cuda")
    w = torch.randn(p, dtype=torch.float64, device="cuda") / p ** 0.5
    z = torch.empty(p, dtype=torch.float64, device="cuda")
    t = torch.empty(n, dtype=torch.float64, device="cuda")
    ss = torch.empty(1, dtype=torch.float64, device="cuda")
    gb = X.numel() * X.element_size() / 1e9
    ms = C.c_float(0)
    out = []
    eng._ck(lib.tpls_op_contract(eng.h, X.data_ptr(), code, n, p, u.data_ptr(), masked, z.data_ptr(), C.byref(ms), 20))
    out.append("contract %.0f" % (gb / ms.value * 1e3))
    eng._ck(lib.tpls_op_project(eng.h, X.data_ptr(), code, n, p, w.data_ptr(), masked, t.data_ptr(), C.byref(ms), 20))
    out.append("project %.0f" % (gb / ms.value * 1e3))
    eng._ck(lib.tpls_op_deflate_contract(eng.h, X.data_ptr(), code, n, p, t.data_ptr(), w.data_ptr(), u.data_ptr(), masked,
                                         z.data_ptr(), ss.data_ptr(), C.byref(ms), 10))
    out.append("deflate %.0f" % (2 * gb / ms.value * 1e3))
    cfg = " ".join(f"{k}={os.environ[k]}" for k in sorted(os.environ) if k.startswith("TPLS_") or k.startswith("TUNE_"))
    print(f"[{cfg}] " + "  ".join(out) + "  GB/s", flush=True)


if __name__ == "__main__":
    main()
