#!/bin/bash
# session-3 call 17: source-level stall samples of the masked row pass (MODE 1) inside a configs[2] fit
O=gpurun_out/r02e17; mkdir -p $O
export TPLS_NO_GRAPH=1
CMD="python tools/config_bench.py --configs 3 --no-cpu"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'rowpass_kernel' --launch-skip 6 -c 1 -f -o $O/prof_rp $CMD > $O/ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $O/ncu.log
ncu -i $O/prof_rp.ncu-rep --page source --csv > $O/source_all.csv 2> /dev/null
python - <<P
import csv
csv.field_size_limit(1 << 30)
rows = list(csv.reader(open("$O/source_all.csv")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; si = h.index("# Samples")
keep = [r for r in rows[hi + 1:] if len(r) > si and r[si] not in ("", "0")]
with open("$O/source_rp.csv", "w", newline="") as f:
    w = csv.writer(f); [w.writerow(r) for r in rows[:hi + 1]]; [w.writerow(r) for r in keep]
print(rows[0][:2], "lines", len(rows) - hi - 1, "with samples", len(keep))
P
rm -f $O/prof_rp.ncu-rep $O/source_all.csv
