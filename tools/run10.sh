python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo rc=$?; tail -3 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo rc=$?; tail -2 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json | cut -c1-600
