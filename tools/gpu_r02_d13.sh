#!/bin/bash
O=gpurun_out/r02d13; mkdir -p $O
L=$PWD/cmtf_pls_b200
echo "== fit_trace, product lib"; TPLS_B200_LIB=$L/libtpls_b200.so TRACE_ITERS=20 timeout 900 python tools/fit_trace.py $O/prod 0 2> $O/t.err | grep -h "contract   \|project"
echo "== fit_trace, product lib, R=10 iters=100"; TPLS_B200_LIB=$L/libtpls_b200.so TRACE_R=10 TRACE_ITERS=100 timeout 900 python tools/fit_trace.py $O/prod10 0 2>> $O/t.err | grep -h "contract   \|project"
echo "== fit_trace, probe lib, R=10 iters=100"; TPLS_B200_LIB=$L/libtpls_b200_probe.so TRACE_R=10 TRACE_ITERS=100 timeout 900 python tools/fit_trace.py $O/probe10 0,128 2>> $O/t.err | grep -h "==\|contract   \|project"
tail -n 3 $O/t.err
