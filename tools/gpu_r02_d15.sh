#!/bin/bash
# session-2 call 15: GPU tests + same-box A/B (HEAD build vs new streaming loops) + all configurations
O=gpurun_out/r02d15; mkdir -p $O
L=$PWD/cmtf_pls_b200
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.txt 2>&1; tail -n 2 $O/pytest.txt
for v in _base "" _base ""; do
  rm -f $O/trace$v.txt
  TPLS_PROFILE_TRACE=$PWD/$O/trace$v.txt TPLS_B200_LIB=$L/libtpls_b200$v.so timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu --quick --no-parity > $O/bench$v.json 2>> $O/bench.err
  python - <<P
import json
d=json.load(open("$O/bench$v.json"))
pc=d["roofline"]["per_class"]
print("lib[$v]", round(d["ms_per_step"],1), "ms  clocks", d["clocks"]["sm_mhz"], {k:round(v["gbs"]) for k,v in pc.items() if v.get("gbs")})
P
  [ -f $O/trace$v.txt ] && python tools/trace_classes.py $O/trace$v.txt | grep "^contract   \|^project\|^deflate"
done
timeout 900 python tools/config_bench.py > $O/configs.jsonl 2> $O/configs.err; tail -n 2 $O/configs.err
python - <<P
import json
for l in open("$O/configs.jsonl"):
    d=json.loads(l); s=d["stream"]; print(d["config"][:40], round(s["ms_device"],3), "ms", round(s["gbs"]), "GB/s", "cov", round(d["covariance"]["ms_device"],3))
P
