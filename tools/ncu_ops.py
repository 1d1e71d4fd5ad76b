"""One launch of every streaming operator variant on a c4-shaped shard (250k x 4096 fp32), for an
`ncu --set full` capture of the kernels the fit-level capture does not reach often: the fused
deflate + contract pass and the NaN-masked variants.

    python tools/ncu_ops.py                       # must exit 0 on its own first
    ncu --set full --clock-control none -k regex:'colpass_kernel|rowpass_kernel|multiproj_kernel|covpass_kernel' -c 14 -f -o gpurun_out/prof_ops python tools/ncu_ops.py
"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from cmtf_pls_b200._core import get_engine  # noqa: E402


def main():
    eng = get_engine(0)
    lib = eng.lib
    n, p = 250_000, 4096
    u = torch.randn(n, dtype=torch.float64, device="cuda")
    t = torch.randn(n, dtype=torch.float64, device="cuda")
    w = torch.randn(p, dtype=torch.float64, device="cuda") / p ** 0.5
    z = torch.empty(p, dtype=torch.float64, device="cuda")
    ss = torch.empty(1, dtype=torch.float64, device="cuda")
    for masked in (0, 1):
        X = torch.randn(n, p, dtype=torch.float32, device="cuda")
        if masked:
            X[torch.rand(n, p, device="cuda") < 0.2] = float("nan")
        torch.cuda.synchronize()
        eng._ck(lib.tpls_op_contract(eng.h, X.data_ptr(), 0, n, p, u.data_ptr(), masked, z.data_ptr(), None, 1))
        eng._ck(lib.tpls_op_project(eng.h, X.data_ptr(), 0, n, p, w.data_ptr(), masked, t.data_ptr(), None, 1))
        eng._ck(lib.tpls_op_deflate_contract(eng.h, X.data_ptr(), 0, n, p, t.data_ptr(), w.data_ptr(), u.data_ptr(), masked,
                                             z.data_ptr(), ss.data_ptr(), None, 1))
        torch.cuda.synchronize()
        del X
    # the single-pass multi-component projection of transform (multiproj.cu) and one cross-covariance pass
    from cmtf_pls_b200 import ctPLS
    X = torch.randn(n, 64, 64, dtype=torch.float32, device="cuda")
    Y = torch.randn(n, 4, dtype=torch.float64, device="cuda")
    est = ctPLS(10, algorithm="covariance")
    est.fit([X[:50_000]], Y[:50_000], max_iter=3)
    est.transform([X])
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
