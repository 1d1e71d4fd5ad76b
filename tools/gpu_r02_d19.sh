#!/bin/bash
O=gpurun_out/r02d19; mkdir -p $O
timeout 300 python tools/resident_probe.py > $O/probe.txt 2>&1; grep -v "^resident" $O/probe.txt; grep "^resident" $O/probe.txt | sed -n '3p;9p'
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_res.txt 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest_res.txt
