MP_TRACE=2 timeout 900 python bench.py --config cfg3 --no-cpu-baseline --steps 2 --warmup 3 --contexts 1 > gpurun_out/bq3.json 2> gpurun_out/bq3.err; grep "mp_trace" gpurun_out/bq3.err | tail -40
