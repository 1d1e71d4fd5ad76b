#!/bin/bash
# ncu evidence for profiles/ (one B200, run under gpurun), round 2.  Each ncu pass runs only after the same command
# exited 0 without it.  The fit is enqueued kernel by kernel (TPLS_NO_GRAPH=1: the same kernels as the graph-launched
# fit, but plain launches that ncu lists one by one).  The reports are summarised ON THE BOX (tools/ncu_summary.py, the
# raw and source pages as CSV) and deleted: gpurun brings back at most 64 MiB.
#   gpurun --timeout 1500 -- 'bash tools/ncu_capture_r02.sh'
set -u
O=gpurun_out/r02ncu
mkdir -p $O
export TPLS_NO_GRAPH=1
CMD="python bench.py --rows 250000 --steps 1 --warmup 0 --no-cpu --quick --no-parity"
$CMD > $O/plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain.log; exit 1; }
MINE='colpass_kernel|rowpass_kernel|covpass_kernel|cov_loop_kernel|rank1_kernel|fold_sets_kernel|reduce_cols_kernel|reduce_q_stop_kernel|row_finish_kernel|finalize_mean_kernel|gather_col_kernel|gram_rows_kernel|lincomb_kernel|multi_dot_kernel|solve_coef_kernel|normalize_q|stop_kernel|reset_ctrl_kernel|transpose_out_kernel|scale_rows_kernel|sum_small_kernel|xchg_kernel|fill_kernel|multiproj|reconstruct_kernel'
# launch list of the FIRST fit of the process (= the timed one: --warmup 0): every kernel of this library
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$MINE" -c 1700 --csv \
    --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1
echo "launch list rc=$?"
python tools/ncu_summary.py launches $O/launches.csv $O/launch_list_summary.csv "TPLS_NO_GRAPH=1 $CMD"
full() {  # name, kernel regex, launches to skip, launches to capture, command...
    local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
    ncu --set full --clock-control none --import-source on -k regex:"$rx" --launch-skip $skip -c $cnt -f -o $O/prof_$name "$@" > $O/ncu_$name.log 2>&1
    echo "full set ($name) rc=$?"
    python tools/ncu_summary.py full $O/prof_$name.ncu-rep $O/full_$name.csv "$*"
    ncu -i $O/prof_$name.ncu-rep --page source --csv > $O/source_$name.csv 2> /dev/null
    rm -f $O/prof_$name.ncu-rep
}
$CMD > $O/plain2.log 2>&1 || { echo "plain run 2 failed"; exit 1; }
# colpass launches of a fit: column statistics (X, X, Y), centre Y, fused centring (X, X), then the contractions
full colpass 'colpass_kernel' 4 4 $CMD
full rowpass 'rowpass_kernel' 2 2 $CMD
full rank1 'rank1_kernel' 2 1 $CMD
python tools/ncu_ops.py > $O/ops_plain.log 2>&1 || { echo "ops plain run failed"; tail -5 $O/ops_plain.log; exit 1; }
full ops 'colpass_kernel|rowpass_kernel' 0 7 python tools/ncu_ops.py
full multiproj 'multiproj_kernel' 0 1 python tools/ncu_ops.py
full covpass 'covpass_kernel' 1 1 python tools/ncu_ops.py
du -sh $O; ls -la $O
