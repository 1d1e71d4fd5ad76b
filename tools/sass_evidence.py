"""Instruction counts per kernel from the built library's SASS (runs without a GPU):
   python tools/sass_evidence.py > profiles/rNN_sass_evidence.csv
UBLKCP = cp.async.bulk (TMA engine, 1-D bulk copy); SYNCS = mbarrier arrive / try_wait."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cmtf_pls_b200", "libtpls_b200.so")
COLS = ["UBLKCP", "SYNCS", "LDS", "STS", "F2F", "DFMA", "DMMA", "DMUL+DADD", "SHFL", "BAR", "LD/LDG", "ST/STG", "LDL/STL"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        base = op.split(".")[0]
        c = counts[cur]
        c["total"] += 1
        if base == "UBLKCP":
            c["UBLKCP"] += 1
        elif base == "SYNCS":
            c["SYNCS"] += 1
        elif base == "DMMA":
            c["DMMA"] += 1
        elif base == "LDS":
            c["LDS"] += 1
        elif base == "STS":
            c["STS"] += 1
        elif base == "F2F":
            c["F2F"] += 1
        elif base == "DFMA":
            c["DFMA"] += 1
        elif base in ("DMUL", "DADD"):
            c["DMUL+DADD"] += 1
        elif base == "SHFL":
            c["SHFL"] += 1
        elif base == "BAR":
            c["BAR"] += 1
        elif base in ("LD", "LDG"):
            c["LD/LDG"] += 1
        elif base in ("ST", "STG"):
            c["ST/STG"] += 1
        elif base in ("LDL", "STL"):
            c["LDL/STL"] += 1
    names = list(counts)
    dem = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines() if names else []
    for n, d in zip(names, dem):
        demangle[n] = d
    print("# SASS evidence: cuobjdump -sass cmtf_pls_b200/libtpls_b200.so (sm_100a), instruction counts per kernel")
    print("# UBLKCP = cp.async.bulk (TMA engine, 1-D bulk copy); SYNCS = mbarrier arrive / try_wait; no tensor-core ops by "
          "design (HBM-bound GEMV-class path)")
    print("kernel,instructions," + ",".join(COLS))
    for n, c in counts.items():
        name = demangle.get(n, n)
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"\((?:int|bool)\)", "", name).replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        name = re.sub(r"\([^()]*\)\s*$", "", name).replace("tpls::", "")
        print('"%s",%d,%s' % (name, c["total"], ",".join(str(c[k]) for k in COLS)))


if __name__ == "__main__":
    sys.exit(main())
