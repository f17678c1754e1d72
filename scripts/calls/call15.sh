set -x
mkdir -p gpurun_out
timeout 900 python scripts/gpu_ab.py 20000 htru PFX_HYBRID=1,PFX_HY_K=32 PFX_HYBRID=1,PFX_HY_K=14 PFX_HYBRID=1,PFX_HY_K=12 PFX_HYBRID=1,PFX_HY_K=8 > gpurun_out/c15_ab.log 2>&1
timeout 900 python scripts/gpu_ab.py 100000 htru PFX_HYBRID=0 PFX_HYBRID=1,PFX_HY_K=32 PFX_HYBRID=1,PFX_HY_K=14 PFX_HYBRID=1,PFX_HY_K=10 >> gpurun_out/c15_ab.log 2>&1
cat gpurun_out/c15_ab.log
