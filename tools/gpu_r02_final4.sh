#!/bin/bash
# session-3 very last call (1 GPU): all GPU tests, smoke and the small configurations on the final code
O=gpurun_out/r02final4; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt; tail -n 3 $O/pytest.txt
timeout 300 python __graft_entry__.py smoke > $O/smoke.txt 2>&1; tail -n 1 $O/smoke.txt
timeout 600 python tools/config_bench.py --configs 1,2,3 --no-cpu > $O/configs.jsonl 2> $O/configs.err
python - <<P
import json
for l in open("$O/configs.jsonl"):
    d=json.loads(l); s=d["stream"]; print(d["config"][:40], round(s["ms_device"],3), "ms", round(s["gbs"]), "GB/s trips", sum(s["trips"]), "launches", s["kernel_launches"], "cov", round(d["covariance"]["ms_device"],3))
P
