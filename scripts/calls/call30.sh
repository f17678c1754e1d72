mkdir -p gpurun_out
timeout 900 python scripts/gpu_step_ab.py 100000 4 - PFX_FILL=1 PFX_FILL=32 PFX_FILL=256 > gpurun_out/c30.log 2>&1
cat gpurun_out/c30.log
