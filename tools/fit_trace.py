"""Per-launch trace of a PROFILED fit of the benchmark workload under probe switches (tools/probe_build.sh, TPLS_DBG):
what each streaming kernel takes inside the fit (power-capped, 1M-row launches) with parts of it switched off.
Results of a fit under a switch are garbage; only the durations count.  Usage: fit_trace.py OUTDIR [dbg,dbg,...]"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)


def main():
    out = sys.argv[1]
    variants = (sys.argv[2] if len(sys.argv) > 2 else "0").split(",")     # "dbg" or "dbg:colpad_kb"
    os.makedirs(out, exist_ok=True)
    import torch
    import bench
    from cmtf_pls_b200 import ctPLS
    rows = int(os.environ.get("TRACE_ROWS", bench.N_TOTAL))
    Xs, Y = bench.make_shard_device(0, rows, torch.device("cuda", 0))
    comps, iters = int(os.environ.get("TRACE_R", 2)), int(os.environ.get("TRACE_ITERS", 12))
    for v in variants:
        dbg, _, pad = v.partition(":")
        os.environ["TPLS_DBG"] = dbg
        os.environ["TPLS_DBG_COLPAD_KB"] = pad or "0"
        path = os.path.join(out, f"trace_dbg{v.replace(':', '_')}.txt")
        if os.path.exists(path):
            os.remove(path)
        est = ctPLS(comps, device=0)
        est.fit(Xs, Y, max_iter=iters)                       # warm-up (graph path)
        os.environ["TPLS_PROFILE_TRACE"] = path
        for _ in range(2):
            est.fit(Xs, Y, max_iter=iters, profile=True)
        _ = est.profile_
        os.environ.pop("TPLS_PROFILE_TRACE")
        print(f"== TPLS_DBG={v}", flush=True)
        print(subprocess.run([sys.executable, os.path.join(HERE, "trace_classes.py"), path], capture_output=True, text=True).stdout,
              flush=True)


if __name__ == "__main__":
    main()
