"""Probe build experiments on the two streaming kernels (tools/probe_build.sh; TPLS_DBG switches, see rowpass.cu).
Times contract / project on the c4 shard shape under each switch and samples clocks + power during long runs."""
import ctypes as C
import json
import os
import sys
import threading
import time

os.environ.setdefault("TPLS_B200_LIB", os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "cmtf_pls_b200",
                                                    "libtpls_b200_probe.so"))
import torch

sys.path.insert(0, ".")
from cmtf_pls_b200._core import get_engine  # noqa: E402
import pynvml  # noqa: E402


def sampler(stop, out):
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    while not stop.is_set():
        out.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM),
                    pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
        time.sleep(0.02)


def main():
    n = int(os.environ.get("TUNE_ROWS", 250_000))
    p = int(os.environ.get("TUNE_P", 4096))
    dt = torch.float32 if os.environ.get("TUNE_DTYPE", "f32") == "f32" else torch.float64
    eng = get_engine(0)
    lib = eng.lib
    code = 0 if dt == torch.float32 else 1
    X = torch.randn(n, p, dtype=dt, device="cuda")
    u = torch.randn(n, dtype=torch.float64, device="cuda")
    w = torch.randn(p, dtype=torch.float64, device="cuda") / p ** 0.5
    z = torch.empty(p, dtype=torch.float64, device="cuda")
    t = torch.empty(n, dtype=torch.float64, device="cuda")
    gb = X.numel() * X.element_size() / 1e9
    ms = C.c_float(0)

    def contract(rep, masked=0):
        eng._ck(lib.tpls_op_contract(eng.h, X.data_ptr(), code, n, p, u.data_ptr(), masked, z.data_ptr(), C.byref(ms), rep))
        return gb / ms.value * 1e3

    def project(rep, masked=0):
        eng._ck(lib.tpls_op_project(eng.h, X.data_ptr(), code, n, p, w.data_ptr(), masked, t.data_ptr(), C.byref(ms), rep))
        return gb / ms.value * 1e3

    rep = int(os.environ.get("PROBE_REP", 40))
    variants = [int(v) for v in os.environ.get("PROBE_DBG", "0,1,2,4,8,16,24,0").split(",")]
    for v in variants:
        os.environ["TPLS_DBG"] = str(v)
        c1, p1 = contract(rep), project(rep)
        c2, p2 = contract(rep), project(rep)
        print(json.dumps({"dbg": v, "contract": [round(c1), round(c2)], "project": [round(p1), round(p2)]}), flush=True)
    os.environ["TPLS_DBG"] = "0"
    # clocks and power under a long run of each kernel
    for name, fn in (("contract", contract), ("project", project), ("contract", contract), ("project", project)):
        stop, out = threading.Event(), []
        th = threading.Thread(target=sampler, args=(stop, out))
        th.start()
        g = fn(3000)
        stop.set()
        th.join()
        out = out[len(out) // 4:]
        med = lambda k: sorted(o[k] for o in out)[len(out) // 2]
        print(json.dumps({"long": name, "gbs": round(g), "sm_mhz": med(0), "mem_mhz": med(1), "power_w": round(med(2)), "samples": len(out)}),
              flush=True)


if __name__ == "__main__":
    main()
