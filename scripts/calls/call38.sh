mkdir -p gpurun_out
for p in 0 1 0 1; do
PFX_PIPE=$p timeout 600 python bench.py --n 1000000 --steps 2 --warmup 3 --no-cpu > gpurun_out/c38_p$p.json 2> gpurun_out/c38_p$p.err
python -c "
import json
d=json.load(open('gpurun_out/c38_p$p.json')); print('pipe $p', 'value %.0f e2e %.0f' % (d['value'], d['e2e']['value']), d['ms_per_step'], d['e2e']['ms_per_step'])"
done
