set -x
N=4
mkdir -p gpurun_out/r02
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555"
$TR bench.py --gpus $N --steps 4 --warmup 3 --no-cpu > gpurun_out/r02/scale_weak_n$N.json 2> gpurun_out/r02/scale_weak_n$N.err
$TR bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --total 800000 > gpurun_out/r02/scale_total800k_n$N.json 2> gpurun_out/r02/scale_total800k_n$N.err
for f in gpurun_out/r02/scale_*_n$N.json; do tail -1 $f | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d['config']['name'], d['n_gpus'], 'value %.0f e2e %.0f ms %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']), d['scaling'], d['config']['candidates_total'])
except Exception as e: print('ERR', e)
"; done
