mkdir -p gpurun_out/r02
timeout 600 python scripts/gpu_configs.py cli > gpurun_out/r02/config0_cli.json 2> gpurun_out/r02/config0_cli.err
cat gpurun_out/r02/config0_cli.json
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu --total 800000 > gpurun_out/r02/scale_total800k_n1.json 2> gpurun_out/r02/scale_total800k_n1.err
python -c "
import json
d=json.load(open('gpurun_out/r02/scale_total800k_n1.json')); print('total800k n1', d['value'], d['e2e']['value'], d['ms_per_step'])"
