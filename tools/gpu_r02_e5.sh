#!/bin/bash
# session-3 call 5: resident loop, rows in L2 first with whole rows in flight, staged kron(w), MASKED as a template parameter
O=gpurun_out/r02e5; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "resident" > $O/pytest_res.txt 2>&1; echo "resident tests rc=$?"; tail -n 3 $O/pytest_res.txt
timeout 300 python tools/resident_probe.py > $O/probe.txt 2>&1; grep -v "^resident" $O/probe.txt; grep "^resident" $O/probe.txt | sed -n '3p;9p'
TPLS_B200_LIB=$PWD/cmtf_pls_b200/libtpls_b200_probe_fine.so timeout 300 python tools/resident_probe.py > $O/probe_fine.txt 2>&1
grep -A1 "^resident" $O/probe_fine.txt | sed -n '5,6p;23,24p'
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_all.txt 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest_all.txt
