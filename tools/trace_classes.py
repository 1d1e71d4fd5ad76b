"""Reads a TPLS_PROFILE_TRACE file (class ms bytes per launch, launch order) and prints, per kernel class and per
POSITION of the launch within a run of the same class (first / second coupled tensor), the mean duration and GB/s."""
import collections
import sys

NAMES = ["colstat", "contract", "project", "deflate_contract", "residual", "rank1", "yside", "other", "nccl", "xchg"]
rows = [l.split() for l in open(sys.argv[1]) if l.strip()]
acc = collections.defaultdict(lambda: [0.0, 0.0, 0])
prev, pos = None, 0
dead = collections.Counter()
for c, ms, b in rows:
    c = int(c)
    if float(b) > 0 and float(b) / (float(ms) * 1e-3 + 1e-12) > 20e12:   # faster than any HBM: the stop flag was up, the kernel returned at once
        dead[c] += 1
        prev = None
        continue
    pos = pos + 1 if c == prev else 0
    prev = c
    a = acc[(c, pos)]
    a[0] += float(ms)
    a[1] += float(b)
    a[2] += 1
for (c, pos), (ms, b, n) in sorted(acc.items()):
    print(f"{NAMES[c]:17s} pos {pos}  launches {n:5d}  mean {ms / n * 1e3:9.1f} us" + (f"  {b / ms / 1e6:8.0f} GB/s" if b else ""))
for c, n in sorted(dead.items()):
    print(f"{NAMES[c]:17s} dead launches (stop flag already up, not counted): {n}")
