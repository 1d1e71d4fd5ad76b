#!/bin/bash
# round-2 GPU call 4 (8 GPUs): the scaling bench, full line (graph loop) + quick A/B lines (PDL, host loop)
set -u
O=gpurun_out/r02c4
mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu"
timeout 900 $B --e2e-steps 3 > $O/bench8.json 2> $O/bench8.err; echo "rc=$?" >> $O/bench8.err
TPLS_PDL=1 timeout 600 $B --quick > $O/bench8_pdl.json 2> $O/bench8_pdl.err; echo "rc=$?" >> $O/bench8_pdl.err
TPLS_NO_GRAPH=1 timeout 600 $B --quick > $O/bench8_hostloop.json 2> $O/bench8_hostloop.err; echo "rc=$?" >> $O/bench8_hostloop.err
for f in $O/bench*.err; do echo "== $f"; tail -n 3 $f; done
