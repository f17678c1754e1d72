set -x
mkdir -p gpurun_out
timeout 600 python scripts/gpu_ab.py 20000 htru PFX_HYBRID=1,PFX_HY_K=32 PFX_SLOTS=1 PFX_SLOTS=1,PFX_HY_K=10 > gpurun_out/c16_ab.log 2>&1
timeout 900 python scripts/gpu_ab.py 100000 htru PFX_HYBRID=1,PFX_HY_K=32 PFX_SLOTS=1 PFX_SLOTS=1,PFX_HY_K=12 PFX_SLOTS=1,PFX_HY_K=8 >> gpurun_out/c16_ab.log 2>&1
timeout 900 python scripts/gpu_ab.py 400000 htru PFX_HYBRID=0 PFX_SLOTS=1 >> gpurun_out/c16_ab.log 2>&1
cat gpurun_out/c16_ab.log
