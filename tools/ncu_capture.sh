#!/bin/bash
# ncu evidence for profiles/ (one B200, run under gpurun): the launch list of one benchmark fit and a full-set
# capture of the streaming and rank-1 kernels.  Each ncu pass runs only after the same command exited 0 without it.
#   gpurun --timeout 900 -- 'bash tools/ncu_capture.sh'
# then, here:  python tools/ncu_summary.py launches gpurun_out/launches_r01c.csv profiles/r01_ncu_launch_list_summary.csv "$CMD"
#              python tools/ncu_summary.py full gpurun_out/prof_r01c.ncu-rep profiles/r01_ncu_full_top_kernels.csv "$CMD"
set -u
CMD="python bench.py --rows 250000 --steps 1 --warmup 1 --no-cpu --e2e-steps 0"
$CMD > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
# launch list: only this library's kernels (namespace tpls), the warm-up fit skipped, about one timed fit captured
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tpls --launch-skip 1900 -c 2100 --csv \
    --log-file gpurun_out/launches_r01c.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 || { echo "plain run 2 failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'colpass_kernel|rowpass_kernel|rank1_kernel' \
    --launch-skip 60 -c 10 -f -o gpurun_out/prof_r01c $CMD > gpurun_out/ncu_full.log 2>&1
echo "full set rc=$?"
