#!/bin/bash
# session-3 call 10: resident loop with advancing pointers (projection, contraction), contiguous rows per row lane
O=gpurun_out/r02e10; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "resident" > $O/pytest_res.txt 2>&1; echo "resident tests rc=$?"; tail -n 3 $O/pytest_res.txt
timeout 300 python tools/resident_probe.py > $O/probe.txt 2>&1; grep -v "^resident" $O/probe.txt; grep "^resident" $O/probe.txt | sed -n '3p;9p'
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_all.txt 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest_all.txt
