"""Times the streaming operators and the rank-1 kernel on one GPU (CUDA events
inside the library, on its own stream).  Prints effective GB/s of X."""

import ctypes as C
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from cmtf_pls_b200._core import get_engine  # noqa: E402


def main():
    eng = get_engine(0)
    lib = eng.lib
    out = []
    shapes = [(250_000, 4096, torch.float32), (125_000, 4096, torch.float64), (1_000_000, 512, torch.float32),
              (2_000_000, 24, torch.float64), (400_000, 2048, torch.float32)]
    if "--only-rank1" in sys.argv:
        shapes = []
    for n, p, dt in shapes:
        code = 0 if dt == torch.float32 else 1
        X = torch.randn(n, p, dtype=dt, device="cuda")
        u = torch.randn(n, dtype=torch.float64, device="cuda")
        w = torch.randn(p, dtype=torch.float64, device="cuda") / p ** 0.5
        z = torch.empty(p, dtype=torch.float64, device="cuda")
        t = torch.empty(n, dtype=torch.float64, device="cuda")
        ss = torch.empty(1, dtype=torch.float64, device="cuda")
        gb = X.numel() * X.element_size() / 1e9
        ms = C.c_float(0)
        for masked in (0, 1):
            eng._ck(lib.tpls_op_contract(eng.h, X.data_ptr(), code, n, p, u.data_ptr(), masked, z.data_ptr(), C.byref(ms), 10))
            out.append(dict(op="contract", n=n, p=p, dtype=str(dt), masked=masked, ms=ms.value, gbs=gb / ms.value * 1e3))
            eng._ck(lib.tpls_op_project(eng.h, X.data_ptr(), code, n, p, w.data_ptr(), masked, t.data_ptr(), C.byref(ms), 10))
            out.append(dict(op="project", n=n, p=p, dtype=str(dt), masked=masked, ms=ms.value, gbs=gb / ms.value * 1e3))
            eng._ck(lib.tpls_op_deflate_contract(eng.h, X.data_ptr(), code, n, p, t.data_ptr(), w.data_ptr(), u.data_ptr(),
                                                 masked, z.data_ptr(), ss.data_ptr(), C.byref(ms), 5))
            out.append(dict(op="deflate_contract", n=n, p=p, dtype=str(dt), masked=masked, ms=ms.value,
                            gbs=2 * gb / ms.value * 1e3))
        # torch reference points: a plain copy and a sum over the same buffer
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        Y = torch.empty_like(X)
        for name, fn, mult in (("torch_copy", lambda: Y.copy_(X), 2), ("torch_sum", lambda: X.sum(), 1)):
            fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            m = e0.elapsed_time(e1) / 5
            out.append(dict(op=name, n=n, p=p, dtype=str(dt), ms=m, gbs=mult * gb / m * 1e3))
        del X, Y
        torch.cuda.empty_cache()
    torch.manual_seed(0)
    for dims, kind in [((64, 64), "easy"), ((32, 16, 8), "easy"), ((64, 32), "easy"), ((24,), "easy"), ((38, 65), "easy"),
                       ((8, 6), "easy"), ((32, 16), "easy"), ((32, 16), "hard"),
                       ((64, 64), "hard"), ((32, 16, 8), "hard"), ((64, 32), "hard"), ((16, 8, 6, 4), "hard")]:
        Z = torch.randn(*dims, dtype=torch.float64, device="cuda")
        if kind == "easy":      # one dominant rank-1 term: large spectral gap, few squarings
            v = [torch.randn(d, dtype=torch.float64, device="cuda") for d in dims]
            if len(dims) == 2:
                Z += 4 * torch.outer(v[0], v[1])
            if len(dims) == 3:
                Z += 4 * torch.einsum("i,j,k->ijk", *v)
        else:                   # twelve comparable rank-1 terms like the covariance of a rank-12 CP data set
            letters = "ijkl"[:len(dims)]
            for r in range(12):
                v = [torch.randn(d, dtype=torch.float64, device="cuda") for d in dims]
                Z += (3.0 - 0.1 * r) * torch.einsum(",".join(letters) + "->" + letters, *v)
        w = torch.zeros(sum(dims), dtype=torch.float64, device="cuda")
        wk = torch.zeros(int(np.prod(dims)), dtype=torch.float64, device="cuda")
        sw = C.c_int(0)
        ms = C.c_float(0)
        eng._ck(lib.tpls_op_rank1(eng.h, Z.data_ptr(), len(dims), (C.c_int * len(dims))(*dims), 1e-8, 0, w.data_ptr(),
                                  wk.data_ptr(), C.byref(sw), C.byref(ms), 20))
        out.append(dict(op="rank1", dims=dims, kind=kind, sweeps=sw.value, us=ms.value * 1e3))
    for r in out:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
