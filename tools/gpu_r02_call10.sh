#!/bin/bash
# round-2 GPU call 10 (4 GPUs): scaling points N=2 and N=4 with the exchange-wait diagnostics
set -u
O=gpurun_out/r02c10
mkdir -p $O
for N in 2 4; do
  B="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2971$N bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --e2e-steps 2"
  timeout 900 $B > $O/bench$N.json 2> $O/bench$N.err; echo "rc=$?" >> $O/bench$N.err
  tail -n 2 $O/bench$N.err
done
