#!/bin/bash
# session-3 final call (1 GPU): all GPU tests, smoke, the judged bench line, every BASELINE configuration, ncu after-captures
set -u
O=gpurun_out/r02final; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt; tail -n 3 $O/pytest.txt
timeout 300 python __graft_entry__.py smoke > $O/smoke.txt 2>&1; tail -n 1 $O/smoke.txt
timeout 900 python bench.py > $O/bench1.json 2> $O/bench1.err; echo "bench rc=$?"; tail -c 600 $O/bench1.json
timeout 900 python tools/config_bench.py --no-cpu > $O/configs.jsonl 2> $O/configs.err
python - <<P
import json
for l in open("$O/configs.jsonl"):
    d=json.loads(l); s=d["stream"]; print(d["config"][:40], round(s["ms_device"],3), "ms", round(s["gbs"]), "GB/s trips", sum(s["trips"]), "launches", s["kernel_launches"], "cov", round(d["covariance"]["ms_device"],3))
P
timeout 300 python tools/resident_probe.py > $O/probe.txt 2>&1; grep -v "^resident" $O/probe.txt; grep "^resident" $O/probe.txt | sed -n '3p;9p'
export TPLS_NO_GRAPH=1
CMD="python tools/config_bench.py --configs 3 --no-cpu"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'rowpass_kernel|colpass_kernel' --launch-skip 12 -c 4 -f -o $O/prof_masked $CMD > $O/ncu_masked.log 2>&1
echo "ncu masked rc=$?"
python tools/ncu_summary.py full $O/prof_masked.ncu-rep $O/full_masked_after.csv "TPLS_NO_GRAPH=1 $CMD"
rm -f $O/prof_masked.ncu-rep
CMD="python tools/resident_ncu.py"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:resident_loop_kernel --launch-skip 5 -c 1 -f -o $O/prof_resident $CMD > $O/ncu_resident.log 2>&1
echo "ncu resident rc=$?"
python tools/ncu_summary.py full $O/prof_resident.ncu-rep $O/full_resident.csv "TPLS_NO_GRAPH=1 $CMD"
rm -f $O/prof_resident.ncu-rep
cut -c1-400 $O/full_masked_after.csv | tail -5
