set -x
mkdir -p gpurun_out
P=$PWD/pulsarfeatureextractor_b200
export PFX_LIB=$P/libpfx_k32.so
python scripts/gpu_gauss_only.py 40000 2 > gpurun_out/c7_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_gauss_rounds_h -s 1 -c 1 -o gpurun_out/r02_rounds_hybrid_k32 -f python scripts/gpu_gauss_only.py 40000 2 > gpurun_out/c7_ncu.log 2>&1
tail -3 gpurun_out/c7_ncu.log
