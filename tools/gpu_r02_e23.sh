#!/bin/bash
# session-3 call 23: per-tensor cache rows (Y and narrow tensors fully cached first)
O=gpurun_out/r02e23; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_parity.txt 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest_parity.txt
timeout 300 python tools/resident_probe.py > $O/probe.txt 2>&1; grep -v "^resident" $O/probe.txt; grep "^resident" $O/probe.txt | sed -n '3p;9p'
