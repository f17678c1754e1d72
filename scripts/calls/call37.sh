mkdir -p gpurun_out
timeout 400 python scripts/gpu_step_ab.py 100000 4 PFX_PIPE=0 PFX_PIPE=1 > gpurun_out/c37.log 2>&1; echo "rc=$?" >> gpurun_out/c37.log
timeout 600 python scripts/gpu_step_ab.py 400000 2 PFX_PIPE=0 PFX_PIPE=1 >> gpurun_out/c37.log 2>&1; echo "rc=$?" >> gpurun_out/c37.log
timeout 900 python bench.py --config target10m --steps 1 --warmup 3 --no-cpu > gpurun_out/c37_bench_t10m.json 2> gpurun_out/c37_bench_t10m.err
cat gpurun_out/c37.log
python -c "
import json
d=json.load(open('gpurun_out/c37_bench_t10m.json')); print('t10m', d['value'], d['e2e']['value'], d['ms_per_step'])"
