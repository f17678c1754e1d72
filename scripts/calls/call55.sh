mkdir -p gpurun_out/r02
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "host_path or large_batch or handover" ) > gpurun_out/c55_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c55_pytest.log
tail -2 gpurun_out/c55_pytest.log
timeout 600 python bench.py --config target10m --steps 1 --warmup 1 --no-cpu > gpurun_out/r02/bench_target10m.json 2> gpurun_out/r02/bench_target10m.err
python -c "
import json
d=json.load(open('gpurun_out/r02/bench_target10m.json')); print('t10m', d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['host_equals_device'])"
