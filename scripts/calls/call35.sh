mkdir -p gpurun_out
timeout 300 python scripts/gpu_step_ab.py 20000 3 - PFX_PIPE=1 > gpurun_out/c35.log 2>&1; echo "rc=$?" >> gpurun_out/c35.log
timeout 400 python scripts/gpu_step_ab.py 100000 4 - PFX_PIPE=1 >> gpurun_out/c35.log 2>&1; echo "rc=$?" >> gpurun_out/c35.log
cat gpurun_out/c35.log
