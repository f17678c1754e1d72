mkdir -p gpurun_out
P=/root/repo/pulsarfeatureextractor_b200
timeout 1200 python scripts/gpu_step_ab.py 100000 4 - PFX_HY_PAD=0,PFX_HY_K=15 PFX_HY_PAD=0,PFX_HY_K=14 PFX_LIB=$P/libpfx_fs4.so > gpurun_out/c31.log 2>&1
cat gpurun_out/c31.log
