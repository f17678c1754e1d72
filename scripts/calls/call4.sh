set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c4_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
tail -15 gpurun_out/c4_pytest.log
python -c "
import json; d=json.load(open('gpurun_out/c4_bench.json')); print(d['value'], d['e2e']['value'], d['stage_ms_serialised'], d['roofline_lyon']['frac'], d['roofline_lyon']['kernel_ms'])"
