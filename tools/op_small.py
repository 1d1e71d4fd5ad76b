"""Streaming operators at the configs[2] size (100k x 2048 fp32, 819 MB): does the pass size explain the in-fit rates?"""
import ctypes as C
import sys

sys.path.insert(0, ".")
import torch
from cmtf_pls_b200._core import get_engine

eng = get_engine(0)
lib = eng.lib
for n, p in ((100_000, 2048), (200_000, 2048), (400_000, 2048), (100_000, 4096)):
    X = torch.randn(n, p, dtype=torch.float32, device="cuda")
    X[torch.rand(n, p, device="cuda") < 0.2] = float("nan")
    u = torch.randn(n, dtype=torch.float64, device="cuda")
    w = torch.randn(p, dtype=torch.float64, device="cuda") / p ** 0.5
    z = torch.empty(p, dtype=torch.float64, device="cuda")
    t = torch.empty(n, dtype=torch.float64, device="cuda")
    gb = X.numel() * 4 / 1e9
    ms = C.c_float(0)
    for masked in (0, 1):
        eng._ck(lib.tpls_op_contract(eng.h, X.data_ptr(), 0, n, p, u.data_ptr(), masked, z.data_ptr(), C.byref(ms), 20))
        c = ms.value
        eng._ck(lib.tpls_op_project(eng.h, X.data_ptr(), 0, n, p, w.data_ptr(), masked, t.data_ptr(), C.byref(ms), 20))
        print(f"{n} x {p} masked={masked}: contract {c * 1e3:7.1f} us {gb / c:7.1f} GB/s   project {ms.value * 1e3:7.1f} us {gb / ms.value:7.1f} GB/s", flush=True)
    del X
