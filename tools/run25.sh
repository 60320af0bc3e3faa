for s in 2 3 5; do
MP_BLOOM_STRIDE=$s MP_BENCH_VERBOSE=1 timeout 300 python bench.py --no-cpu-baseline --steps 4 > gpurun_out/bq_s$s.json 2> gpurun_out/bq_s$s.err; echo "stride $s"; grep "loop R" gpurun_out/bq_s$s.err | cut -c1-120
done
