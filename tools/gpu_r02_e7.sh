#!/bin/bash
# session-3 call 7: one-warp matrix path of the rank-1 step: parity tests, opbench, resident probe
O=gpurun_out/r02e7; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_all.txt 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest_all.txt
TPLS_RANK1_STAMPS=1 timeout 300 python tools/opbench.py --only-rank1 > $O/rank1.txt 2>&1; grep "^{" $O/rank1.txt
timeout 300 python tools/resident_probe.py > $O/probe.txt 2>&1; grep -v "^resident" $O/probe.txt; grep "^resident" $O/probe.txt | sed -n '3p;9p'
