set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/c36_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c36_pytest.log
tail -8 gpurun_out/c36_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/c36_bench.json 2> gpurun_out/c36_bench.err
timeout 900 python bench.py --config target10m --steps 1 --warmup 3 --no-cpu > gpurun_out/c36_bench_t10m.json 2> gpurun_out/c36_bench_t10m.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c36_smoke.log 2>&1; tail -2 gpurun_out/c36_smoke.log
python -c "
import json
for f in ['c36_bench','c36_bench_t10m']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['e2e']['value'], d['ms_per_step'], d.get('stage_ms_serialised'), d.get('host_equals_device'))
    except Exception as e: print(f, 'ERR', e)
"
