mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "survey or pfd_family or edge or large_batch or reserve" ) > gpurun_out/c41_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c41_pytest.log
tail -3 gpurun_out/c41_pytest.log
python bench.py --config survey --steps 2 --warmup 3 --no-cpu > gpurun_out/c41_survey.json 2> gpurun_out/c41_survey.err
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/c41_htru.json 2> gpurun_out/c41_htru.err
python -c "
import json
for f in ['c41_survey','c41_htru']:
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d.get('stage_ms_serialised'))"
