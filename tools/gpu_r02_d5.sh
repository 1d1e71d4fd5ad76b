#!/bin/bash
O=gpurun_out/r02d5; mkdir -p $O
L=$PWD/cmtf_pls_b200
TPLS_B200_LIB=$L/libtpls_b200_probe.so timeout 900 python tools/fit_trace.py $O/rpi1 16,20,16,20 > $O/rpi1.txt 2> $O/rpi1.err
grep -h "==\|contract   \|project" $O/rpi1.txt; tail -n 3 $O/rpi1.err
