set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_dmfit_exact|k_hist_lane_u8|k_sinusoid' -s 3 -c 3 -o gpurun_out/r02_side_kernels -f python scripts/gpu_bates_once.py 40000 > gpurun_out/c28_ncu.log 2>&1
tail -3 gpurun_out/c28_ncu.log
