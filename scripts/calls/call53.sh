mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_subband -s 1 -c 1 -o gpurun_out/r02_subband_pfd -f python scripts/gpu_configs.py pfd > gpurun_out/c53_ncu.log 2>&1
tail -2 gpurun_out/c53_ncu.log
