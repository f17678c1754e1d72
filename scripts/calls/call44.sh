mkdir -p gpurun_out
P=/root/repo/pulsarfeatureextractor_b200
timeout 900 python scripts/gpu_step_ab.py 100000 4 - PFX_LIB=$P/libpfx_w10.so,PFX_HY_WARPS=10 PFX_LIB=$P/libpfx_w12.so,PFX_HY_WARPS=12 PFX_LIB=$P/libpfx_w16.so,PFX_HY_WARPS=16 PFX_LIB=$P/libpfx_w12.so > gpurun_out/c44.log 2>&1
cat gpurun_out/c44.log
