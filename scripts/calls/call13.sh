set -x
mkdir -p gpurun_out
export PFX_WAVE=1
python scripts/gpu_gauss_only.py 40000 2 > gpurun_out/c13_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_gauss_wave -s 1 -c 1 -o gpurun_out/r02_wave_b -f python scripts/gpu_gauss_only.py 40000 2 > gpurun_out/c13_ncu.log 2>&1
tail -2 gpurun_out/c13_ncu.log
