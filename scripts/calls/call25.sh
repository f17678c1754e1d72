set -x
mkdir -p gpurun_out
export PFX_FINAL_SLOTS=1
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/c25_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c25_pytest.log
tail -12 gpurun_out/c25_pytest.log
ncu --set full --clock-control none --import-source on -k regex:k_gauss_final_s -s 1 -c 1 -o gpurun_out/r02_final_s_a -f python scripts/gpu_gauss_only.py 40000 2 > gpurun_out/c25_ncu.log 2>&1
tail -2 gpurun_out/c25_ncu.log
