mkdir -p gpurun_out/r02
timeout 600 python scripts/gpu_configs.py cli > gpurun_out/r02/config0_cli.json 2> gpurun_out/r02/config0_cli.err
cat gpurun_out/r02/config0_cli.json; tail -3 gpurun_out/r02/config0_cli.err
