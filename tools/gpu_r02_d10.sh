#!/bin/bash
O=gpurun_out/r02d10; mkdir -p $O
L=$PWD/cmtf_pls_b200
TPLS_B200_LIB=$L/libtpls_b200_probe.so timeout 900 python tools/fit_trace.py $O/rpi1 0,2,4,0 > $O/rpi1.txt 2> $O/rpi1.err
TPLS_B200_LIB=$L/libtpls_b200_probe_rpi2.so timeout 900 python tools/fit_trace.py $O/rpi2 0,0 > $O/rpi2.txt 2> $O/rpi2.err
grep -h "==\|contract   \|project" $O/rpi1.txt; echo; grep -h "==\|contract   \|project" $O/rpi2.txt; tail -n 3 $O/rpi1.err
