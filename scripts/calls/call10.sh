set -x
mkdir -p gpurun_out
python scripts/gpu_gauss_only.py 40000 2 > gpurun_out/c10_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'k_gauss_final|k_hist_lane_u8|k_gauss_rounds' -s 3 -c 3 -o gpurun_out/r02_gauss_chain -f python scripts/gpu_gauss_only.py 40000 2 > gpurun_out/c10_ncu.log 2>&1
tail -2 gpurun_out/c10_ncu.log
ncu --clock-control none -k regex:'k_gauss_final|k_gauss_rounds' -s 2 -c 2 --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,gpu__time_duration.sum --csv --log-file gpurun_out/r02_fp64_ops.csv python scripts/gpu_gauss_only.py 100000 2 > gpurun_out/c10_ncu2.log 2>&1
cat gpurun_out/r02_fp64_ops.csv | tail -12
