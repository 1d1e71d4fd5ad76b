#!/bin/bash
# session-3: the judged bench line on the final code (small configurations measured first)
O=gpurun_out/r02final5; mkdir -p $O
timeout 900 python bench.py > $O/bench1.json 2> $O/bench1.err; echo "bench rc=$?"
python - <<P
import json
d = json.loads(open("$O/bench1.json").read().strip().splitlines()[-1])
print("bench:", d["ms_per_step"], "ms", round(d["value"]), d["unit"], "frac", d["config"]["fraction_of_hbm_peak"], "e2e", round(d["e2e"]["s_per_step"], 3), "s", "clocks", d["clocks"])
print("small", d["config"]["small_configs"])
print("roofline", {k: d["roofline"][k] for k in ("kernel", "achieved", "peak", "frac", "traffic", "share_of_step")})
P
