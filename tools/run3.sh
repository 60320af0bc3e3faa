set -x
export MP_BENCH_VERBOSE=1
MP_TRACE=1 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo rc=$?; grep -v "^\[mp_trace\]" gpurun_out/bench_full.err | tail -8; tail -16 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
ls -la /tmp/mpbench/*/ | head -20
