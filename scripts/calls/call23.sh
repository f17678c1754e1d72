set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c23_t100k.csv python scripts/gpu_gauss_only.py 100000 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c23_t400k.csv python scripts/gpu_gauss_only.py 400000 2 > /dev/null 2>&1
grep -h "k_gauss\|k_hist" gpurun_out/c23_t100k.csv gpurun_out/c23_t400k.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120
ncu --set full --clock-control none --import-source on -k regex:k_gauss_final -s 1 -c 1 -o gpurun_out/r02_final_a -f python scripts/gpu_gauss_only.py 40000 2 > gpurun_out/c23_ncu.log 2>&1
tail -2 gpurun_out/c23_ncu.log
