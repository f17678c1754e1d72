set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/c5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c5_pytest.log
python scripts/gpu_parity_table.py > gpurun_out/c5_parity.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err
PFX_DM_EXACT=0 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/c5_bench_lane.json 2> gpurun_out/c5_bench_lane.err
python bench.py --config pfd --steps 3 --warmup 3 --no-cpu > gpurun_out/c5_bench_pfd.json 2> gpurun_out/c5_bench_pfd.err
PFX_DM_EXACT=0 python bench.py --config pfd --steps 3 --warmup 3 --no-cpu > gpurun_out/c5_bench_pfd_lane.json 2> gpurun_out/c5_bench_pfd_lane.err
tail -15 gpurun_out/c5_pytest.log
python -c "
import json
for f in ['c5_bench','c5_bench_lane','c5_bench_pfd','c5_bench_pfd_lane']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['e2e']['value'], d.get('stage_ms_serialised'))
    except Exception as e: print(f, 'ERR', e)
"
