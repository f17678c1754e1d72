mkdir -p gpurun_out
timeout 900 python scripts/gpu_step_ab.py 100000 4 - PFX_HY_WARPS=4,PFX_HY_K=14 PFX_HY_WARPS=4,PFX_HY_K=13 PFX_HY_WARPS=2,PFX_HY_K=14 > gpurun_out/c43.log 2>&1
cat gpurun_out/c43.log
