set -x
mkdir -p gpurun_out
timeout 600 python scripts/gpu_ab.py 20000 htru PFX_FINAL_SLOTS=0 PFX_FINAL_SLOTS=1 > gpurun_out/c24_ab.log 2>&1
timeout 900 python scripts/gpu_ab.py 100000 htru PFX_FINAL_SLOTS=0 PFX_FINAL_SLOTS=1 >> gpurun_out/c24_ab.log 2>&1
cat gpurun_out/c24_ab.log
