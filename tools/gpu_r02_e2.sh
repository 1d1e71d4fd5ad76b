#!/bin/bash
# session-3 call 2: resident loop with the shared-memory row cache: resident tests, probe with and without the cache, all GPU tests
O=gpurun_out/r02e2; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "resident" > $O/pytest_res.txt 2>&1; echo "resident tests rc=$?"; tail -n 5 $O/pytest_res.txt
timeout 300 python tools/resident_probe.py > $O/probe.txt 2>&1; grep -v "^resident" $O/probe.txt; grep "^resident" $O/probe.txt | sed -n '3p;9p'
TPLS_RESIDENT_CACHE=0 timeout 300 python tools/resident_probe.py > $O/probe_nocache.txt 2>&1; grep -v "^resident" $O/probe_nocache.txt; grep "^resident" $O/probe_nocache.txt | sed -n '3p;9p'
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_all.txt 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest_all.txt
