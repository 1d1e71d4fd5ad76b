#!/bin/bash
# session-3 call 6: ncu full set + source page of one resident-loop launch (configs[1], component 1 of the second fit)
O=gpurun_out/r02e6; mkdir -p $O
export TPLS_NO_GRAPH=1
CMD="python tools/resident_ncu.py"
$CMD > $O/plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain.log; exit 1; }
cat $O/plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resident_loop_kernel --launch-skip 5 -c 1 -f -o $O/prof_resident $CMD > $O/ncu_resident.log 2>&1
echo "ncu rc=$?"; tail -3 $O/ncu_resident.log
python tools/ncu_summary.py full $O/prof_resident.ncu-rep $O/full_resident.csv "TPLS_NO_GRAPH=1 $CMD"
ncu -i $O/prof_resident.ncu-rep --page source --csv > $O/source_resident_all.csv 2> /dev/null
python - <<P
import csv, sys
csv.field_size_limit(1 << 30)
rows = list(csv.reader(open("$O/source_resident_all.csv")))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
si = hdr.index("# Samples")
keep = [r for r in rows[hdr_i + 1:] if len(r) > si and r[si] not in ("", "0")]
with open("$O/source_resident.csv", "w", newline="") as f:
    w = csv.writer(f); [w.writerow(r) for r in rows[:hdr_i + 1]]; [w.writerow(r) for r in keep]
print("source lines", len(rows) - hdr_i - 1, "with samples", len(keep))
P
ncu -i $O/prof_resident.ncu-rep --page details --csv > $O/details_resident.csv 2> /dev/null
rm -f $O/prof_resident.ncu-rep $O/source_resident_all.csv
ls -la $O
