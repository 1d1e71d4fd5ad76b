#!/bin/bash
# round-2 GPU call 5 (1 GPU): PDL on by default, cluster cov loop, 4-chain DMMA, bounce-buffered reconstruction
set -u
O=gpurun_out/r02c5
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_default.txt 2>&1; echo "rc=$?" >> $O/pytest_default.txt
TPLS_PDL=0 timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_nopdl.txt 2>&1; echo "rc=$?" >> $O/pytest_nopdl.txt
TPLS_RANK1_STAMPS=1 timeout 600 python tools/opbench.py --only-rank1 > $O/opbench_rank1.jsonl 2> $O/opbench_rank1.err
timeout 900 python tools/config_bench.py --no-cpu > $O/configs.jsonl 2> $O/configs.err
timeout 900 python bench.py --steps 3 --warmup 3 --e2e-steps 2 > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
for f in $O/pytest_*.txt; do echo "== $f"; tail -n 6 $f; done
tail -n 3 $O/bench.err
