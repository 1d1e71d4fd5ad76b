#!/bin/bash
# round-2 GPU call 6 (1 GPU): staged epilogue inputs of the row pass; same-box A/B of the in-fit kernel variants
set -u
O=gpurun_out/r02c6
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_default.txt 2>&1; echo "rc=$?" >> $O/pytest_default.txt
TPLS_NO_STAGE_AUX=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q > $O/pytest_nostage.txt 2>&1; echo "rc=$?" >> $O/pytest_nostage.txt
Q="python bench.py --steps 3 --warmup 3 --quick --no-parity --no-cpu"
timeout 600 $Q > $O/ab_default.json 2> $O/ab_default.err
TPLS_B200_LIB=$PWD/cmtf_pls_b200/libtpls_b200_sschain1.so timeout 600 $Q > $O/ab_sschain1.json 2> $O/ab_sschain1.err
TPLS_NO_STAGE_AUX=1 timeout 600 $Q > $O/ab_nostage.json 2> $O/ab_nostage.err
timeout 600 $Q > $O/ab_default2.json 2> $O/ab_default2.err
timeout 900 python tools/config_bench.py --no-cpu --configs 1,2,3,5 > $O/configs.jsonl 2> $O/configs.err
for f in $O/pytest_*.txt; do echo "== $f"; tail -n 6 $f; done
