mkdir -p gpurun_out
run() { timeout 600 python bench.py --n 2097152 --steps 1 --warmup 1 --no-cpu > gpurun_out/c54.json 2> gpurun_out/c54.err; python -c "
import json
d=json.load(open('gpurun_out/c54.json')); print('$1', 'value %.0f e2e %.0f' % (d['value'], d['e2e']['value']), d['e2e']['ms_per_step'])"; }
PFX_CHUNK=262144 run c256k
PFX_CHUNK=524288 run c512k
PFX_CHUNK=1048576 run c1m
PFX_CHUNK=262144 run c256k
