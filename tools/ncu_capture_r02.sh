#!/bin/bash
# ncu evidence for profiles/ (one B200, run under gpurun), round 2.  Each ncu pass runs only after the same command
# exited 0 without it.  The fit is enqueued kernel by kernel (TPLS_NO_GRAPH=1: the same kernels as the graph-launched
# fit, but plain launches that ncu lists one by one).
#   gpurun --timeout 1500 -- 'bash tools/ncu_capture_r02.sh'
# then, here:  python tools/ncu_summary.py launches gpurun_out/r02ncu/launches.csv profiles/r02_ncu_launch_list_summary.csv "$CMD"
#              python tools/ncu_summary.py full gpurun_out/r02ncu/prof_fit.ncu-rep profiles/r02_ncu_full_top_kernels.csv "$CMD"
#              python tools/ncu_summary.py full gpurun_out/r02ncu/prof_ops.ncu-rep profiles/r02_ncu_full_op_variants.csv "python tools/ncu_ops.py"
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
$CMD > $O/plain2.log 2>&1 || { echo "plain run 2 failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'colpass_kernel|rowpass_kernel|rank1_kernel' \
    --launch-skip 40 -c 12 -f -o $O/prof_fit $CMD > $O/ncu_full.log 2>&1
echo "full set (fit) rc=$?"
# (TPLS_NO_GRAPH stays set: the covariance fit of ncu_ops.py is enqueued kernel by kernel too)
python tools/ncu_ops.py > $O/ops_plain.log 2>&1 || { echo "ops plain run failed"; tail -5 $O/ops_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'colpass_kernel|rowpass_kernel' \
    -c 7 -f -o $O/prof_ops python tools/ncu_ops.py > $O/ncu_ops.log 2>&1
echo "full set (ops) rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'multiproj_kernel' \
    -c 1 -f -o $O/prof_multiproj python tools/ncu_ops.py > $O/ncu_multiproj.log 2>&1
echo "full set (multiproj) rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'covpass_kernel' \
    --launch-skip 1 -c 1 -f -o $O/prof_covpass python tools/ncu_ops.py > $O/ncu_covpass.log 2>&1
echo "full set (covpass) rc=$?"
ls -la $O
