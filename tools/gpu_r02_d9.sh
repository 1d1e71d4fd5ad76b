#!/bin/bash
O=gpurun_out/r02d9; mkdir -p $O
L=$PWD/cmtf_pls_b200
TPLS_B200_LIB=$L/libtpls_b200_probe.so TUNE_ROWS=1000000 PROBE_DBG=0,64,16,80,0,64 PROBE_REP=10 timeout 900 python tools/probe_streams.py > $O/p.txt 2> $O/p.err
cat $O/p.txt; tail -n 3 $O/p.err
