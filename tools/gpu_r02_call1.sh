#!/bin/bash
# round-2 GPU call 1: graph WHILE smoke test, GPU test suite (graph / host loop / PDL), config bench, short bench
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c1
mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1
timeout 60 tools/micro/cond_while > $O/cond_while.txt 2>&1; echo "cond_while rc=$?" >> $O/cond_while.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_graph.txt 2>&1; echo "rc=$?" >> $O/pytest_graph.txt
TPLS_NO_GRAPH=1 timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_hostloop.txt 2>&1; echo "rc=$?" >> $O/pytest_hostloop.txt
TPLS_PDL=1 timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_graph_pdl.txt 2>&1; echo "rc=$?" >> $O/pytest_graph_pdl.txt
timeout 900 python tools/config_bench.py --no-cpu > $O/configs_graph.jsonl 2> $O/configs_graph.err
TPLS_PDL=1 timeout 900 python tools/config_bench.py --no-cpu --configs 1,2,3,5 > $O/configs_graph_pdl.jsonl 2> $O/configs_graph_pdl.err
TPLS_NO_GRAPH=1 timeout 900 python tools/config_bench.py --no-cpu --configs 1,2,3,5 > $O/configs_hostloop.jsonl 2> $O/configs_hostloop.err
timeout 900 python bench.py --steps 2 --warmup 3 --e2e-steps 1 > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
tail -3 $O/pytest_graph.txt $O/pytest_hostloop.txt $O/pytest_graph_pdl.txt $O/cond_while.txt
