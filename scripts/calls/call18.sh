set -x
mkdir -p gpurun_out
timeout 900 python scripts/gpu_ab.py 100000 htru PFX_HYBRID=1,PFX_HY_K=32 PFX_SLOTS=1 PFX_HYBRID=0 > gpurun_out/c18_ab.log 2>&1
timeout 900 python scripts/gpu_ab.py 400000 htru PFX_SLOTS=1 >> gpurun_out/c18_ab.log 2>&1
cat gpurun_out/c18_ab.log
