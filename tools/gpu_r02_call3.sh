#!/bin/bash
# round-2 GPU call 3 (1 GPU): full suite with the DMMA rank-1 / single-pass transform / robust regression, opbench, configs, bench
set -u
O=gpurun_out/r02c3
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_graph.txt 2>&1; echo "rc=$?" >> $O/pytest_graph.txt
TPLS_NO_GRAPH=1 timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_hostloop.txt 2>&1; echo "rc=$?" >> $O/pytest_hostloop.txt
TPLS_RANK1_STAMPS=1 timeout 600 python tools/opbench.py > $O/opbench.jsonl 2> $O/opbench.err
timeout 900 python tools/config_bench.py --no-cpu > $O/configs_graph.jsonl 2> $O/configs_graph.err
timeout 900 python bench.py --steps 3 --warmup 3 --e2e-steps 2 > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
for f in $O/pytest_*.txt; do echo "== $f"; tail -n 6 $f; done
tail -n 3 $O/bench.err
