#!/bin/bash
# round-2 GPU call 11 (2 GPUs): exchange kernel with parallel announcements / polls, 8-chain folds, narrow folds
set -u
O=gpurun_out/r02c11
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
tail -n 4 $O/pytest.txt
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --e2e-steps 2"
timeout 900 $B > $O/bench2.json 2> $O/bench2.err; echo "rc=$?" >> $O/bench2.err
tail -n 2 $O/bench2.err
timeout 600 python tools/config_bench.py --no-cpu --configs 1,2,3,5 > $O/configs.jsonl 2> $O/configs.err
