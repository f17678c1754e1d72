set -x
mkdir -p gpurun_out
timeout 1200 python scripts/gpu_ab.py 100000 htru PFX_LMPAR_BUDGET=10 PFX_LMPAR_BUDGET=1 PFX_LMPAR_BUDGET=2 PFX_LMPAR_BUDGET=3 > gpurun_out/c26_ab.log 2>&1
cat gpurun_out/c26_ab.log
