#!/bin/bash
# session-3 call 18: row pass without a branch per column group (rows that do not fill the CTA): tests, configs[2], configs[1] streaming
O=gpurun_out/r02e19; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_all.txt 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest_all.txt
timeout 600 python tools/config_bench.py --configs 3,5 --no-cpu > $O/configs.jsonl 2> $O/configs.err
python - <<P
import json
for l in open("$O/configs.jsonl"):
    d=json.loads(l); s=d["stream"]; print(d["config"][:40], round(s["ms_device"],3), "ms", round(s["gbs"]), "GB/s trips", sum(s["trips"]), "launches", s["kernel_launches"], "cov", round(d["covariance"]["ms_device"],3)); print(s.get("profile_ms"))
P
