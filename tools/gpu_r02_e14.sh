#!/bin/bash
# session-3 call 14: streaming operators at the configs[2] pass size
O=gpurun_out/r02e14; mkdir -p $O
timeout 300 python tools/op_small.py > $O/op_small.txt 2>&1; cat $O/op_small.txt
