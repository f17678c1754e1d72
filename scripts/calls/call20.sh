set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/c20_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c20_pytest.log
tail -15 gpurun_out/c20_pytest.log
python scripts/gpu_parity_table.py > gpurun_out/c20_parity.log 2>&1
tail -5 gpurun_out/c20_parity.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/c20_bench.json 2> gpurun_out/c20_bench.err
python bench.py --config pfd --steps 3 --warmup 3 --no-cpu > gpurun_out/c20_bench_pfd.json 2> gpurun_out/c20_bench_pfd.err
python bench.py --config survey --steps 2 --warmup 3 --no-cpu > gpurun_out/c20_bench_survey.json 2> gpurun_out/c20_bench_survey.err
python -c "
import json
for f in ['c20_bench','c20_bench_pfd','c20_bench_survey']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['e2e']['value'], d.get('stage_ms_serialised'))
    except Exception as e: print(f, 'ERR', e)
"
