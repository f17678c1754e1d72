set -x
mkdir -p gpurun_out/r02
python bench.py --steps 5 --warmup 3 > gpurun_out/r02/bench_htru100k.json 2> gpurun_out/r02/bench_htru100k.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02/bench_reference.json 2> gpurun_out/r02/bench_reference.err
python bench.py --config pfd --steps 3 --warmup 3 --no-cpu > gpurun_out/r02/bench_pfd.json 2> gpurun_out/r02/bench_pfd.err
python bench.py --config survey --steps 2 --warmup 3 --no-cpu > gpurun_out/r02/bench_survey.json 2> gpurun_out/r02/bench_survey.err
python bench.py --config lyon10m --steps 5 --warmup 3 --no-cpu > gpurun_out/r02/bench_lyon10m.json 2> gpurun_out/r02/bench_lyon10m.err
python bench.py --config target10m --steps 1 --warmup 3 --no-cpu > gpurun_out/r02/bench_target10m.json 2> gpurun_out/r02/bench_target10m.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02/ncu_launch.log 2>&1
for f in gpurun_out/r02/bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print(d.get('config',{}).get('workload','')[:40], d.get('value'), d.get('e2e',{}).get('value'), d.get('ms_per_step'))
except Exception as e: print('ERR', e)
"; done
