#!/bin/bash
# session-3 call 24: the tail of a component (regression, Y deflation) inside the resident launch
O=gpurun_out/r02e24; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest.txt
timeout 300 python tools/resident_probe.py > $O/probe.txt 2>&1; grep -v "^resident" $O/probe.txt
timeout 600 python tools/config_bench.py --configs 1,2 --no-cpu > $O/configs.jsonl 2> $O/configs.err
python - <<P
import json
for l in open("$O/configs.jsonl"):
    d=json.loads(l); s=d["stream"]; print(d["config"][:40], round(s["ms_device"],3), "ms", round(s["gbs"]), "GB/s trips", sum(s["trips"]), "launches", s["kernel_launches"], "cov", round(d["covariance"]["ms_device"],3))
P
