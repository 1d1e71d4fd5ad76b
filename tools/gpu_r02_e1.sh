#!/bin/bash
# session-3 call 1: rank-1 phase stamps on the current build
O=gpurun_out/r02e1; mkdir -p $O
TPLS_RANK1_STAMPS=1 timeout 300 python tools/opbench.py --only-rank1 > $O/rank1.txt 2>&1; cat $O/rank1.txt
