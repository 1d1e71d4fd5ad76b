#!/bin/bash
# round-2 session-2 call 1: streaming-kernel probe (TPLS_DBG switches) + GPU tests of the refactored tile walk
O=gpurun_out/r02d1; mkdir -p $O
timeout 600 python tools/probe_streams.py > $O/probe.jsonl 2> $O/probe.err; echo "rc=$?" >> $O/probe.err
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.txt 2>&1; tail -n 3 $O/pytest.txt
cat $O/probe.jsonl; tail -n 3 $O/probe.err
