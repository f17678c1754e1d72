set -x
mkdir -p gpurun_out
nproc > gpurun_out/c1_nproc.txt
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
python scripts/gpu_parity_table.py > gpurun_out/c1_parity.log 2>&1
python scripts/gpu_subbatch.py 100000 3 > gpurun_out/c1_subbatch.log 2>&1
python scripts/gpu_subbatch.py 400000 2 >> gpurun_out/c1_subbatch.log 2>&1
PFX_LIB=$PWD/pulsarfeatureextractor_b200/libpfx_prof.so python scripts/gpu_phase_profile.py 100000 > gpurun_out/r02_phase_profile_baseline.txt 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
tail -3 gpurun_out/c1_pytest.log; cat gpurun_out/c1_subbatch.log; cat gpurun_out/r02_phase_profile_baseline.txt
