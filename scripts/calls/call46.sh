mkdir -p gpurun_out
run() { python bench.py --config pfd --steps 2 --warmup 3 --no-cpu > gpurun_out/c46.json 2> gpurun_out/c46.err; python -c "
import json
d=json.load(open('gpurun_out/c46.json')); print('$1', 'value %.0f e2e %.0f ms %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']), {k: round(v,1) for k,v in d['stage_ms_serialised'].items()})"; }
run base
PFX_SUB_GLOBAL=1 PFX_SUB_BLOCKS=2 run global2
PFX_SUB_GLOBAL=1 PFX_SUB_BLOCKS=4 run global4
