set -x
mkdir -p gpurun_out
P=$PWD/pulsarfeatureextractor_b200
python scripts/gpu_variants.py 100000 gaussian libpfx.so libpfx_k32.so > gpurun_out/c6_variants.log 2>&1
PFX_HYBRID=0 python scripts/gpu_variants.py 100000 gaussian libpfx.so >> gpurun_out/c6_variants.log 2>&1
python scripts/gpu_variants.py 400000 gaussian libpfx.so libpfx_k32.so >> gpurun_out/c6_variants.log 2>&1
PFX_HYBRID=0 python scripts/gpu_variants.py 400000 gaussian libpfx.so >> gpurun_out/c6_variants.log 2>&1
cat gpurun_out/c6_variants.log
( time timeout 600 python -m pytest tests -m gpu -q -x ) > gpurun_out/c6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c6_pytest.log
tail -15 gpurun_out/c6_pytest.log
PFX_LIB=$P/libpfx_prof.so python scripts/gpu_phase_profile.py 100000 > gpurun_out/r02_phase_profile_hybrid_k16.txt 2>&1
PFX_LIB=$P/libpfx_k32prof.so python scripts/gpu_phase_profile.py 100000 > gpurun_out/r02_phase_profile_hybrid_k32.txt 2>&1
cat gpurun_out/r02_phase_profile_hybrid_k16.txt gpurun_out/r02_phase_profile_hybrid_k32.txt
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err
python -c "
import json
d=json.load(open('gpurun_out/c6_bench.json')); print(d['value'], d['e2e']['value'], d.get('stage_ms_serialised'))"
