#!/bin/bash
# session-3 call 4: fine stamps inside the resident loop (probe build)
O=gpurun_out/r02e4; mkdir -p $O
TPLS_B200_LIB=$PWD/cmtf_pls_b200/libtpls_b200_probe_fine.so timeout 300 python tools/resident_probe.py > $O/probe.txt 2>&1
grep -v "^resident\|cycles per" $O/probe.txt; grep -A1 "^resident" $O/probe.txt | sed -n '5,6p;23,24p'
