#!/bin/bash
# round-2 GPU call 2 (2 GPUs): full GPU suite incl. the sharded tests, 2-rank bench (graph, graph+PDL, host loop, NCCL)
set -u
O=gpurun_out/r02c2
mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_graph.txt 2>&1; echo "rc=$?" >> $O/pytest_graph.txt
TPLS_PDL=1 timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > $O/pytest_multi_pdl.txt 2>&1; echo "rc=$?" >> $O/pytest_multi_pdl.txt
TPLS_NO_GRAPH=1 timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > $O/pytest_multi_hostloop.txt 2>&1; echo "rc=$?" >> $O/pytest_multi_hostloop.txt
TPLS_NO_XCHG=1 timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > $O/pytest_multi_nccl.txt 2>&1; echo "rc=$?" >> $O/pytest_multi_nccl.txt
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu"
timeout 900 $B > $O/bench2.json 2> $O/bench2.err; echo "rc=$?" >> $O/bench2.err
TPLS_PDL=1 timeout 900 $B > $O/bench2_pdl.json 2> $O/bench2_pdl.err; echo "rc=$?" >> $O/bench2_pdl.err
TPLS_NO_GRAPH=1 timeout 900 $B > $O/bench2_hostloop.json 2> $O/bench2_hostloop.err; echo "rc=$?" >> $O/bench2_hostloop.err
timeout 900 python bench.py --steps 2 --warmup 3 --e2e-steps 1 > $O/bench1.json 2> $O/bench1.err; echo "rc=$?" >> $O/bench1.err
for f in $O/pytest_*.txt; do echo "== $f"; tail -n 4 $f; done
for f in $O/bench*.err; do echo "== $f"; tail -n 3 $f; done
