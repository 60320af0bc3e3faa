MP_BENCH_VERBOSE=1 timeout 400 python bench.py --no-cpu-baseline --steps 6 > gpurun_out/bq.json 2> gpurun_out/bq.err; grep "loop R" gpurun_out/bq.err | cut -c1-100
