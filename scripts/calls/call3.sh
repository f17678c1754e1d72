set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c3_pytest.log
P=$PWD/pulsarfeatureextractor_b200
python scripts/gpu_variants.py 100000 gaussian libpfx.so libpfx_nocoop.so > gpurun_out/c3_variants.log 2>&1
python scripts/gpu_variants.py 100000 all libpfx.so libpfx_nocoop.so >> gpurun_out/c3_variants.log 2>&1
PFX_LIB=$P/libpfx_prof.so python scripts/gpu_phase_profile.py 100000 > gpurun_out/r02_phase_profile_coop.txt 2>&1
PFX_LIB=$P/libpfx_nocoopprof.so python scripts/gpu_phase_profile.py 100000 > gpurun_out/r02_phase_profile_nocoop.txt 2>&1
python scripts/gpu_parity_table.py > gpurun_out/c3_parity.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
tail -5 gpurun_out/c3_pytest.log; cat gpurun_out/c3_variants.log; cat gpurun_out/r02_phase_profile_coop.txt gpurun_out/r02_phase_profile_nocoop.txt; python -c "
import json; d=json.load(open('gpurun_out/c3_bench.json')); print(d['value'], d['e2e']['value'], d['stage_ms_serialised'], d['roofline_lyon']['frac'], d['roofline_lyon']['kernel_ms'])"
