set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/c22_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c22_pytest.log
tail -8 gpurun_out/c22_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/c22_bench.json 2> gpurun_out/c22_bench.err
python -c "
import json
for f in ['c22_bench']:
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['e2e']['value'], d.get('stage_ms_serialised'))
"
