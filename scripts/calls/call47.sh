mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "pfd or survey or group_entry or edge" ) > gpurun_out/c47_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c47_pytest.log
tail -3 gpurun_out/c47_pytest.log
python bench.py --config pfd --steps 3 --warmup 3 --no-cpu > gpurun_out/c47_pfd.json 2> gpurun_out/c47_pfd.err
python -c "
import json
d=json.load(open('gpurun_out/c47_pfd.json')); print('pfd', 'value %.0f e2e %.0f ms %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']), {k: round(v,1) for k,v in d['stage_ms_serialised'].items()})"
