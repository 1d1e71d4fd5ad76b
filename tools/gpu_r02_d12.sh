#!/bin/bash
O=gpurun_out/r02d12; mkdir -p $O
L=$PWD/cmtf_pls_b200
TPLS_B200_LIB=$L/libtpls_b200_probe.so TRACE_ITERS=20 timeout 900 python tools/fit_trace.py $O/t 0,1,128,0,128 > $O/t.txt 2> $O/t.err
grep -h "==\|contract   \|project" $O/t.txt; tail -n 3 $O/t.err
