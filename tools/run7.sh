python -m pytest tests -m gpu -x -q 2>&1 | tail -12
export MP_BENCH_VERBOSE=1
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo rc=$?; grep "loop A" gpurun_out/bench_full.err | cut -c1-900; cat gpurun_out/bench_full.json | cut -c1-1800
