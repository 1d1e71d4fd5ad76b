#!/bin/bash
O=gpurun_out/r02d7; mkdir -p $O
L=$PWD/cmtf_pls_b200
TPLS_B200_LIB=$L/libtpls_b200_probe.so timeout 900 python tools/fit_trace.py $O/t 16,48,0,32,0,32 > $O/t.txt 2> $O/t.err
grep -h "==\|contract   \|project" $O/t.txt; tail -n 3 $O/t.err
