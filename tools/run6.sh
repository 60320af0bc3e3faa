ARGS="--ref-mbp 50 --pairs-per-step 131072 --steps 1 --warmup 1 --no-cpu-baseline"
python bench.py $ARGS > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_dp" -s 2 -c 1 -o gpurun_out/prof_dp_v0 -f python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1; echo ncu rc=$?
tail -2 gpurun_out/ncu_full.log | cut -c1-300
