set -x
mkdir -p gpurun_out/r02
ncu --clock-control none -k regex:'k_gauss_final_s|k_gauss_rounds_s' -s 2 -c 2 --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file gpurun_out/r02_fp64_ops.csv python scripts/gpu_gauss_only.py 100000 2 > gpurun_out/c32_ncu2.log 2>&1
python scripts/ncu_gauss_json.py gpurun_out/r02_fp64_ops.csv 100000 > gpurun_out/c32_json.log 2>&1
cp profiles/r02_fp64_flops.json profiles/r02_gauss_summary.json gpurun_out/r02/
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r02/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02/pytest_gpu.log
tail -6 gpurun_out/r02/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02/bench_htru100k.json 2> gpurun_out/r02/bench_htru100k.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02/bench_reference.json 2> gpurun_out/r02/bench_reference.err
python bench.py --config pfd --steps 3 --warmup 3 --no-cpu > gpurun_out/r02/bench_pfd.json 2> gpurun_out/r02/bench_pfd.err
python bench.py --config survey --steps 2 --warmup 3 --no-cpu > gpurun_out/r02/bench_survey.json 2> gpurun_out/r02/bench_survey.err
python bench.py --config lyon10m --steps 5 --warmup 3 --no-cpu > gpurun_out/r02/bench_lyon10m.json 2> gpurun_out/r02/bench_lyon10m.err
python bench.py --config target10m --steps 1 --warmup 3 --no-cpu > gpurun_out/r02/bench_target10m.json 2> gpurun_out/r02/bench_target10m.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_gauss_final_s|k_gauss_rounds_s' -s 2 -c 2 -o gpurun_out/r02_gauss_slots_final_build -f python scripts/gpu_gauss_only.py 40000 2 > gpurun_out/c32_ncu3.log 2>&1
for f in gpurun_out/r02/bench_*.json; do python -c "
import json,sys
try:
    d=json.load(open('$f')); print('$f', d.get('value'), d.get('e2e',{}).get('value'), d.get('ms_per_step'), d.get('roofline_fp64',{}).get('frac'))
except Exception as e: print('$f ERR', e)
"; done
