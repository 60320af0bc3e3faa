set -x
ARGS="--steps 1 --warmup 1 --no-cpu-baseline"
python bench.py $ARGS > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_mmp|k_dp" -c 2 -o gpurun_out/prof_v0 -f python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1; echo ncu rc=$?
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
