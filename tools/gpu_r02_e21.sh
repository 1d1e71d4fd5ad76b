#!/bin/bash
# session-3 call 21: rank-1 step on every CTA of the resident loop; goldens through both loops; A/B against the first-L-CTAs form
O=gpurun_out/r02e21; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_all.txt 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest_all.txt
timeout 300 python tools/resident_probe.py > $O/probe.txt 2>&1; grep -v "^resident" $O/probe.txt; grep "^resident" $O/probe.txt | sed -n '3p;9p'
TPLS_RESIDENT_R1_ALL=0 timeout 300 python tools/resident_probe.py > $O/probe_first.txt 2>&1; grep -v "^resident" $O/probe_first.txt; grep "^resident" $O/probe_first.txt | sed -n '3p;9p'
timeout 600 python tools/config_bench.py --configs 1,2 --no-cpu > $O/configs.jsonl 2> $O/configs.err
python - <<P
import json
for l in open("$O/configs.jsonl"):
    d=json.loads(l); s=d["stream"]; print(d["config"][:40], round(s["ms_device"],3), "ms", round(s["gbs"]), "GB/s trips", sum(s["trips"]), "launches", s["kernel_launches"], "cov", round(d["covariance"]["ms_device"],3))
P
