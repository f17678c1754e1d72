set -x
mkdir -p gpurun_out
export PFX_SLOTS=1
python scripts/gpu_gauss_only.py 40000 2 > gpurun_out/c19_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_gauss_rounds_s -s 1 -c 1 -o gpurun_out/r02_slots_b -f python scripts/gpu_gauss_only.py 40000 2 > gpurun_out/c19_ncu.log 2>&1
tail -2 gpurun_out/c19_ncu.log
