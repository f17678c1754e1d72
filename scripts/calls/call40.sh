set -x
mkdir -p gpurun_out/r02
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r02/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02/pytest_gpu.log
tail -6 gpurun_out/r02/pytest_gpu.log
timeout 600 python scripts/gpu_parity_table.py > gpurun_out/r02/parity_table.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02/smoke.log 2>&1; tail -1 gpurun_out/r02/smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02/bench_htru100k.json 2> gpurun_out/r02/bench_htru100k.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02/bench_reference.json 2> gpurun_out/r02/bench_reference.err
python bench.py --config pfd --steps 3 --warmup 3 --no-cpu > gpurun_out/r02/bench_pfd.json 2> gpurun_out/r02/bench_pfd.err
python bench.py --config survey --steps 2 --warmup 3 --no-cpu > gpurun_out/r02/bench_survey.json 2> gpurun_out/r02/bench_survey.err
python bench.py --config lyon10m --steps 5 --warmup 3 --no-cpu > gpurun_out/r02/bench_lyon10m.json 2> gpurun_out/r02/bench_lyon10m.err
python bench.py --config target10m --steps 1 --warmup 3 --no-cpu > gpurun_out/r02/bench_target10m.json 2> gpurun_out/r02/bench_target10m.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02/ncu_launch.log 2>&1; echo "ncu rc=$?"
for f in gpurun_out/r02/bench_*.json; do python -c "
import json,sys
try:
    d=json.load(open('$f')); print('$f', d.get('value'), d.get('e2e',{}).get('value'), d.get('ms_per_step'), d.get('roofline_fp64',{}).get('frac'))
except Exception as e: print('$f ERR', e)
"; done
