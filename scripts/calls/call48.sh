mkdir -p gpurun_out
timeout 1200 python scripts/gpu_step_ab.py 100000 4 - PFX_LMPAR_BUDGET8=3 PFX_LMPAR_BUDGET8=1 PFX_LMPAR_BUDGET=1,PFX_LMPAR_BUDGET8=2 PFX_FS_K=11 PFX_FS_K=9 PFX_PIPE_HOLD=4 PFX_PIPE_HOLD=20 > gpurun_out/c48.log 2>&1
cat gpurun_out/c48.log
