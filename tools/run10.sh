python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo rc=$?; tail -2 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo rc=$?; cut -c1-400 gpurun_out/bench_ref.json
python bench.py --steps 4 --warmup 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 500 --csv --log-file gpurun_out/launches_v4.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; echo ncu rc=$?
