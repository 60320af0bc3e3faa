ARGS="--steps 1 --warmup 1 --no-cpu-baseline --contexts 1"
python bench.py $ARGS > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_mmp" -s 1 -c 1 -o gpurun_out/prof_mmp -f python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1; echo ncu rc=$?
