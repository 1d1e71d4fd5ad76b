#!/bin/bash
# round-2 GPU call 8 (1 GPU): row pass with early-released staged epilogue inputs; same-box A/B against the unstaged path
set -u
O=gpurun_out/r02c8
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest_default.txt 2>&1; echo "rc=$?" >> $O/pytest_default.txt
Q="python bench.py --steps 3 --warmup 3 --quick --no-parity --no-cpu"
timeout 600 $Q > $O/ab_default.json 2> $O/ab_default.err
TPLS_NO_STAGE_AUX=1 timeout 600 $Q > $O/ab_nostage.json 2> $O/ab_nostage.err
timeout 600 $Q > $O/ab_default2.json 2> $O/ab_default2.err
TPLS_NO_STAGE_AUX=1 timeout 600 $Q > $O/ab_nostage2.json 2> $O/ab_nostage2.err
timeout 900 python tools/config_bench.py --no-cpu --configs 2,3,5 > $O/configs.jsonl 2> $O/configs.err
TPLS_NO_STAGE_AUX=1 timeout 900 python tools/config_bench.py --no-cpu --configs 2,3,5 > $O/configs_nostage.jsonl 2> $O/configs_nostage.err
for f in $O/pytest_*.txt; do echo "== $f"; tail -n 6 $f; done
