#!/bin/bash
O=gpurun_out/r02d14; mkdir -p $O
L=$PWD/cmtf_pls_b200
TPLS_B200_LIB=$L/libtpls_b200_probe.so TRACE_ITERS=20 timeout 900 python tools/fit_trace.py $O/t 0,1,128,129,0,1,128,129 2> $O/t.err | grep -h "==\|project"
tail -n 3 $O/t.err
