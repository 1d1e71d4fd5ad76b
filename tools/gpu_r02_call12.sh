#!/bin/bash
# round-2 GPU call 12 (8 GPUs): final scaling point with the exchange diagnostics
set -u
O=gpurun_out/r02c12
mkdir -p $O
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu --e2e-steps 3"
timeout 900 $B > $O/bench8.json 2> $O/bench8.err; echo "rc=$?" >> $O/bench8.err
tail -n 2 $O/bench8.err
