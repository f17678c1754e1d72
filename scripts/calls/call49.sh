set -x
mkdir -p gpurun_out/r02
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02/pytest_gpu.log
tail -8 gpurun_out/r02/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02/smoke.log 2>&1; tail -1 gpurun_out/r02/smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02/bench_htru100k.json 2> gpurun_out/r02/bench_htru100k.err
python -c "
import json
d=json.load(open('gpurun_out/r02/bench_htru100k.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['host_equals_device'], d['cpu_baseline']['value'], d['roofline_fp64']['frac'], d['roofline_lyon']['frac'])"
