#!/bin/bash
# session-2 call 16: first run of the resident trip loop: smoke, GPU tests (resident on / off), small configurations
O=gpurun_out/r02d16; mkdir -p $O
timeout 120 python __graft_entry__.py smoke > $O/smoke.txt 2>&1; echo "smoke rc=$?"; tail -n 3 $O/smoke.txt
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_res.txt 2>&1; echo "pytest(resident auto) rc=$?"; tail -n 12 $O/pytest_res.txt
TPLS_RESIDENT=0 timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_nores.txt 2>&1; echo "pytest(resident off) rc=$?"; tail -n 2 $O/pytest_nores.txt
timeout 600 python tools/config_bench.py > $O/configs.jsonl 2> $O/configs.err; tail -n 2 $O/configs.err
python - <<P
import json
for l in open("$O/configs.jsonl"):
    d=json.loads(l); s=d["stream"]; print(d["config"][:40], round(s["ms_device"],3), "ms", round(s["gbs"]), "GB/s trips", sum(s["trips"]), "launches", s["kernel_launches"], "cov", round(d["covariance"]["ms_device"],3))
P
