set -x
mkdir -p gpurun_out/r02
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r02/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02/pytest_gpu.log
tail -6 gpurun_out/r02/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02/smoke.log 2>&1; tail -1 gpurun_out/r02/smoke.log
timeout 300 ncu --clock-control none -k regex:k_lyon8 -s 2 -c 1 --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --csv --log-file gpurun_out/r02_lyon_metrics.csv python scripts/gpu_lyon.py 4000000 > gpurun_out/c45_lyon.log 2>&1
tail -3 gpurun_out/r02_lyon_metrics.csv | cut -c1-300
python bench.py --steps 5 --warmup 3 > gpurun_out/r02/bench_htru100k.json 2> gpurun_out/r02/bench_htru100k.err
python -c "
import json
d=json.load(open('gpurun_out/r02/bench_htru100k.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['host_equals_device'], d['cpu_baseline']['value'])"
