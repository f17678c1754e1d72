set -x
mkdir -p gpurun_out
P=/root/repo/pulsarfeatureextractor_b200
timeout 900 python scripts/gpu_ab.py 100000 htru PFX_LIB=$P/libpfx_prev.so PFX_LIB=$P/libpfx.so > gpurun_out/c27_ab.log 2>&1
cat gpurun_out/c27_ab.log
