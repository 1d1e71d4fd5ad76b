#!/bin/bash
# session-2 call 3: per-launch trace of the profiled fit (which projection / contraction is the slow one)
O=gpurun_out/r02d3; mkdir -p $O; rm -f $O/trace.txt
TPLS_PROFILE_TRACE=$PWD/$O/trace.txt timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --quick --no-parity > $O/bench.json 2> $O/bench.err
python tools/trace_classes.py $O/trace.txt | tee $O/trace_summary.txt
head -c 600 $O/bench.json; tail -n 2 $O/bench.err
