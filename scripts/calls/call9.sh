set -x
mkdir -p gpurun_out
python scripts/gpu_ab.py 100000 htru PFX_HIST_U8=0 PFX_HIST_U8=1 > gpurun_out/c9_ab.log 2>&1
python scripts/gpu_ab.py 20000 superb64 PFX_HIST_U8=0 PFX_HIST_U8=1 >> gpurun_out/c9_ab.log 2>&1
cat gpurun_out/c9_ab.log
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/c9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c9_pytest.log
tail -5 gpurun_out/c9_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err
python -c "
import json
d=json.load(open('gpurun_out/c9_bench.json')); print(d['value'], d['e2e']['value'], d.get('stage_ms_serialised'))"
