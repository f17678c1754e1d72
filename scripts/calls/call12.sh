set -x
mkdir -p gpurun_out
timeout 600 python scripts/gpu_ab.py 20000 htru PFX_HYBRID=1 PFX_WAVE=1 > gpurun_out/c12_ab.log 2>&1
cat gpurun_out/c12_ab.log
timeout 600 python scripts/gpu_ab.py 100000 htru PFX_WAVE=0 PFX_WAVE=1 > gpurun_out/c12_ab2.log 2>&1
cat gpurun_out/c12_ab2.log
