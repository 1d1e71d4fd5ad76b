#!/bin/bash
# ncu evidence for profiles/ (one B200, run under gpurun): the launch list of one benchmark fit and a full-set
# capture of the streaming and rank-1 kernels.  Each ncu pass runs only after the same command exited 0 without it.
#   gpurun --timeout 900 -- 'bash tools/ncu_capture.sh'
# then, here:  python tools/ncu_summary.py launches gpurun_out/launches_r01c.csv profiles/r01_ncu_launch_list_summary.csv "$CMD"
#              python tools/ncu_summary.py full gpurun_out/prof_r01c.ncu-rep profiles/r01_ncu_full_top_kernels.csv "$CMD"
set -u
CMD="python bench.py --rows 250000 --steps 1 --warmup 1 --no-cpu --e2e-steps 0"
$CMD > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
# launch list: only this library's kernels (ncu matches the base name, without the tpls:: namespace), the warm-up fit
# (1899 launches) skipped, the timed fit captured
MINE='colpass_kernel|rowpass_kernel|covpass_kernel|cov_loop_kernel|rank1_kernel|reduce_cols_kernel|reduce_q_stop_kernel|row_finish_kernel|finalize_mean_kernel|gather_col_kernel|gram_rows_kernel|lincomb_kernel|multi_dot_kernel|solve_coef_kernel|normalize_q|stop_kernel|reset_ctrl_kernel|transpose_out_kernel|scale_rows_kernel|sum_small_kernel|xchg_kernel|fill_kernel'
if [ "${NCU_SKIP_LIST:-0}" != "1" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$MINE" --launch-skip 1899 -c 1899 --csv \
    --log-file gpurun_out/launches_r01c.csv $CMD > gpurun_out/ncu_launches.log 2>&1
fi
echo "launch list rc=$?"
[ "${NCU_SKIP_FULL:-0}" = "1" ] && exit 0
$CMD > gpurun_out/ncu_plain2.log 2>&1 || { echo "plain run 2 failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'colpass_kernel|rowpass_kernel|rank1_kernel' \
    --launch-skip 60 -c 10 -f -o gpurun_out/prof_r01c $CMD > gpurun_out/ncu_full.log 2>&1
echo "full set rc=$?"
