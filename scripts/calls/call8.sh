set -x
mkdir -p gpurun_out
P=$PWD/pulsarfeatureextractor_b200
python scripts/gpu_variants.py 100000 gaussian libpfx.so libpfx_k32.so > gpurun_out/c8_variants.log 2>&1
PFX_HYBRID=0 python scripts/gpu_variants.py 100000 gaussian libpfx.so >> gpurun_out/c8_variants.log 2>&1
cat gpurun_out/c8_variants.log
PFX_LIB=$P/libpfx_prof.so python scripts/gpu_phase_profile.py 100000 > gpurun_out/r02_phase_profile_hybrid_k16.txt 2>&1
PFX_LIB=$P/libpfx_k32prof.so python scripts/gpu_phase_profile.py 100000 > gpurun_out/r02_phase_profile_hybrid_k32.txt 2>&1
head -14 gpurun_out/r02_phase_profile_hybrid_k16.txt; head -14 gpurun_out/r02_phase_profile_hybrid_k32.txt
export PFX_LIB=$P/libpfx_k32.so
ncu --set full --clock-control none --import-source on -k regex:k_gauss_rounds_h -s 1 -c 1 -o gpurun_out/r02_rounds_hybrid_k32b -f python scripts/gpu_gauss_only.py 40000 2 > gpurun_out/c8_ncu.log 2>&1
tail -2 gpurun_out/c8_ncu.log
