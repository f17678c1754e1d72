mkdir -p gpurun_out
run() { timeout 600 python bench.py --n 1000000 --steps 2 --warmup 3 --no-cpu > gpurun_out/c39.json 2> gpurun_out/c39.err; python -c "
import json
d=json.load(open('gpurun_out/c39.json')); print('$1', 'value %.0f e2e %.0f' % (d['value'], d['e2e']['value']), d['ms_per_step'], d['e2e']['ms_per_step'])"; }
PFX_PIPE=0 run pipe0
PFX_PIPE=1 PFX_PIPE_PRIO=1 run pipe1_hi
PFX_PIPE=1 PFX_PIPE_PRIO=0 run pipe1_lo
PFX_PIPE=0 run pipe0
PFX_PIPE=1 PFX_PIPE_PRIO=1 run pipe1_hi
timeout 400 python scripts/gpu_step_ab.py 100000 4 PFX_PIPE=0 PFX_PIPE=1,PFX_PIPE_PRIO=1 PFX_PIPE=1,PFX_PIPE_PRIO=0
