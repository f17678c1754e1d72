mkdir -p gpurun_out
timeout 600 python scripts/gpu_dm_outliers.py 40000 1000 > gpurun_out/c29.log 2>&1
cat gpurun_out/c29.log
