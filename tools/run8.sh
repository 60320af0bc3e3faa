ARGS="--steps 2 --warmup 1 --no-cpu-baseline"
MP_TRACE=1 python bench.py $ARGS > gpurun_out/b.json 2> gpurun_out/b.err; grep mp_trace gpurun_out/b.err | tail -22
python bench.py $ARGS > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/launches_full.csv python bench.py $ARGS > gpurun_out/ncu.log 2>&1; echo ncu rc=$?
