#!/bin/bash
# session-2 call 4: in-fit traces of the streaming kernels under the probe switches (RPI=1 and RPI=2 builds)
O=gpurun_out/r02d4; mkdir -p $O
L=$PWD/cmtf_pls_b200
TPLS_B200_LIB=$L/libtpls_b200_probe.so timeout 900 python tools/fit_trace.py $O/rpi1 0,1,2,4,16,0 > $O/rpi1.txt 2> $O/rpi1.err
TPLS_B200_LIB=$L/libtpls_b200_probe_rpi2.so timeout 600 python tools/fit_trace.py $O/rpi2 0,4,0 > $O/rpi2.txt 2> $O/rpi2.err
grep -h "==\|contract   \|project" $O/rpi1.txt; echo; grep -h "==\|contract   \|project" $O/rpi2.txt; tail -n 3 $O/rpi1.err
