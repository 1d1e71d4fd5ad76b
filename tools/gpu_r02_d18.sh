#!/bin/bash
O=gpurun_out/r02d18; mkdir -p $O
timeout 300 python tools/resident_probe.py > $O/probe.txt 2>&1; tail -n 30 $O/probe.txt
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_res.txt 2>&1; echo "pytest rc=$?"; tail -n 5 $O/pytest_res.txt
