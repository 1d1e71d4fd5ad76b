#!/bin/bash
# session-2 call 17: resident trip loop with loads in flight: GPU tests + small configurations
O=gpurun_out/r02d17; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_res.txt 2>&1; echo "pytest rc=$?"; tail -n 5 $O/pytest_res.txt
timeout 600 python tools/config_bench.py > $O/configs.jsonl 2> $O/configs.err; tail -n 2 $O/configs.err
python - <<P
import json
for l in open("$O/configs.jsonl"):
    d=json.loads(l); s=d["stream"]; print(d["config"][:40], round(s["ms_device"],3), "ms", round(s["gbs"]), "GB/s trips", sum(s["trips"]), "launches", s["kernel_launches"], "cov", round(d["covariance"]["ms_device"],3))
P
