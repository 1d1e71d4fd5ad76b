#!/bin/bash
# session-3 last call (1 GPU): all GPU tests, smoke, the judged bench line, every BASELINE configuration -- on the final code
set -u
O=gpurun_out/r02final2; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt; tail -n 3 $O/pytest.txt
timeout 300 python __graft_entry__.py smoke > $O/smoke.txt 2>&1; tail -n 1 $O/smoke.txt
timeout 900 python bench.py > $O/bench1.json 2> $O/bench1.err; echo "bench rc=$?"
timeout 900 python tools/config_bench.py --no-cpu > $O/configs.jsonl 2> $O/configs.err
python - <<P
import json
d = json.loads(open("$O/bench1.json").read().strip().splitlines()[-1])
print("bench:", d["ms_per_step"], "ms", round(d["value"]), d["unit"], "e2e", round(d["e2e"]["s_per_step"], 3), "s", "small", d["config"]["small_configs"])
for l in open("$O/configs.jsonl"):
    d=json.loads(l); s=d["stream"]; print(d["config"][:40], round(s["ms_device"],3), "ms", round(s["gbs"]), "GB/s trips", sum(s["trips"]), "launches", s["kernel_launches"], "cov", round(d["covariance"]["ms_device"],3))
P
