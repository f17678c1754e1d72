mkdir -p gpurun_out
timeout 900 python scripts/gpu_step_ab.py 100000 4 - PFX_SIN_PER_SM=2,PFX_HY_K=13 PFX_SIN_PER_SM=2 PFX_SIN_PER_SM=4,PFX_HY_K=13 PFX_HY_K=13 > gpurun_out/c42.log 2>&1
cat gpurun_out/c42.log
