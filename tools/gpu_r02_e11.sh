#!/bin/bash
# session-3 call 11: resident/streaming crossover sweep + the BASELINE configurations
O=gpurun_out/r02e11; mkdir -p $O
timeout 600 python tools/resident_sweep.py > $O/sweep.txt 2>&1; cat $O/sweep.txt
timeout 900 python tools/config_bench.py > $O/configs.jsonl 2> $O/configs.err; tail -n 2 $O/configs.err
python - <<P
import json
for l in open("$O/configs.jsonl"):
    d=json.loads(l); s=d["stream"]; print(d["config"][:40], round(s["ms_device"],3), "ms", round(s["gbs"]), "GB/s trips", sum(s["trips"]), "launches", s["kernel_launches"], "cov", round(d["covariance"]["ms_device"],3))
P
