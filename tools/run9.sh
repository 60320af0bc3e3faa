ARGS="--steps 1 --warmup 1 --no-cpu-baseline --contexts 1"
python bench.py $ARGS > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_mmp|k_dp_fill|k_dp_tb" -s 4 -c 3 -o gpurun_out/prof_v3 -f python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1; echo ncu rc=$?
