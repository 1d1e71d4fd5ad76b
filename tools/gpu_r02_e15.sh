#!/bin/bash
# session-3 call 15: ncu full set of the masked row pass (MODE 1) and the masked contraction inside a configs[2] fit
O=gpurun_out/r02e15; mkdir -p $O
export TPLS_NO_GRAPH=1
CMD="python tools/config_bench.py --configs 3 --no-cpu"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'rowpass_kernel|colpass_kernel' --launch-skip 12 -c 4 -f -o $O/prof_masked $CMD > $O/ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $O/ncu.log
python tools/ncu_summary.py full $O/prof_masked.ncu-rep $O/full_masked.csv "TPLS_NO_GRAPH=1 $CMD"
cat $O/full_masked.csv
ncu -i $O/prof_masked.ncu-rep --page source --csv > $O/source_all.csv 2> /dev/null
python - <<P
import csv
csv.field_size_limit(1 << 30)
txt = open("$O/source_all.csv").read().split('"Kernel Name"')
for blk in txt[1:]:
    rows = list(csv.reader(('"Kernel Name"' + blk).splitlines()))
    name = rows[0][1][:70]
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hi]
    si = h.index("# Samples")
    st = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = {}
    for r in rows[hi + 1:]:
        if len(r) <= si: continue
        for i in st:
            try: tot[h[i]] = tot.get(h[i], 0) + int(r[i] or 0)
            except ValueError: pass
    s = sum(tot.values()) or 1
    print(name, {k[6:]: round(100 * v / s, 1) for k, v in sorted(tot.items(), key=lambda x: -x[1])[:7]})
P
rm -f $O/prof_masked.ncu-rep $O/source_all.csv
