#!/bin/bash
# session-2 call 11: GPU tests of the new streaming loops + same-box A/B of the row-pass variants inside the bench fit
O=gpurun_out/r02d11; mkdir -p $O
L=$PWD/cmtf_pls_b200
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.txt 2>&1; tail -n 2 $O/pytest.txt
for v in _base "" _probe_lazy _probe_rpi2 _probe_rpi2lazy _base ""; do
  rm -f $O/trace$v.txt
  TPLS_PROFILE_TRACE=$PWD/$O/trace$v.txt TPLS_B200_LIB=$L/libtpls_b200$v.so timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu --quick --no-parity > $O/bench$v.json 2>> $O/bench.err
  python - <<P
import json
d=json.load(open("$O/bench$v.json"))
print("lib[$v]", round(d["ms_per_step"],1), "ms  clocks", d["clocks"]["sm_mhz"])
P
  [ -f $O/trace$v.txt ] && python tools/trace_classes.py $O/trace$v.txt | grep "^contract   \|^project\|^deflate"
done
