timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:k_dp_fill -c 1 -o gpurun_out/fill_cw python bench.py --profile-step --no-cpu-baseline > gpurun_out/ncu_fill.log 2>&1
tail -3 gpurun_out/ncu_fill.log
