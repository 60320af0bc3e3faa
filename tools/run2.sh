set -x
export MP_BENCH_VERBOSE=1
ARGS="--ref-mbp 50 --pairs-per-step 262144 --steps 2 --warmup 1 --cpu-sample-pairs 50000"
MP_TRACE=1 python bench.py $ARGS > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo rc=$?; grep -v "^\[" gpurun_out/bench_small.err | tail -5; tail -40 gpurun_out/bench_small.err; cat gpurun_out/bench_small.json
python bench.py $ARGS --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_small.csv python bench.py $ARGS --no-cpu-baseline > gpurun_out/ncu.log 2>&1; echo ncu rc=$?
