set -x
mkdir -p gpurun_out/r02
nvidia-smi -L
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "two_gpus" ) > gpurun_out/r02/pytest_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02/pytest_2gpu.log
tail -4 gpurun_out/r02/pytest_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
$TR bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu > gpurun_out/r02/scale_weak_n2.json 2> gpurun_out/r02/scale_weak_n2.err
$TR bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --total 800000 > gpurun_out/r02/scale_total800k_n2.json 2> gpurun_out/r02/scale_total800k_n2.err
$TR bench.py --gpus 2 --config target10m --steps 1 --warmup 3 --no-cpu > gpurun_out/r02/scale_target10m_n2.json 2> gpurun_out/r02/scale_target10m_n2.err
for f in gpurun_out/r02/scale_*_n2.json; do python -c "
import json
try:
    d=json.load(open('$f')); print('$f', d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['scaling'])
except Exception as e: print('$f ERR', e)
"; done
