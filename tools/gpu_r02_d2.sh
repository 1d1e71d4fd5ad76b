#!/bin/bash
# session-2 call 2: same-box A/B of the row pass (stage released before the arithmetic of a tile's last rows; RPI=1 / 2)
O=gpurun_out/r02d2; mkdir -p $O
L=cmtf_pls_b200
for v in base probe probe_rpi2 base probe probe_rpi2; do
  TPLS_B200_LIB=$PWD/$L/libtpls_b200_$v.so PROBE_DBG=0,0 timeout 300 python tools/probe_streams.py > $O/probe_$v.jsonl 2>> $O/probe.err
  echo "== $v"; cat $O/probe_$v.jsonl
done
for v in base probe base probe; do
  TPLS_B200_LIB=$PWD/$L/libtpls_b200_$v.so timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --quick --no-parity > $O/bench_$v.json 2>> $O/bench.err
  python - <<P
import json
d=json.load(open("$O/bench_$v.json"))
pc=d["roofline"]["per_class"]
print("$v", d["ms_per_step"], {k:round(v["gbs"]) for k,v in pc.items() if "gbs" in v and v["gbs"]})
P
done
