#!/bin/bash
# session-3 call 20 (2 GPUs): all GPU tests incl. the sharded ones + a short sharded bench (parity gate inside)
set -u
O=gpurun_out/r02e20
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
tail -n 4 $O/pytest.txt
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --quick"
timeout 900 $B > $O/bench2.json 2> $O/bench2.err; echo "rc=$?" >> $O/bench2.err
tail -n 2 $O/bench2.err
python - <<P
import json
d = json.loads(open("$O/bench2.json").read().strip().splitlines()[-1])
print("2 GPUs:", d["ms_per_step"], "ms", d["value"], d["unit"], "parity", d.get("parity_check"), "trips", d["config"]["trips_total"])
P
