export MP_BENCH_VERBOSE=1
for g in 32 64 128; do echo "== L2 fetch $g"; MP_L2_FETCH=$g python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 >/dev/null | grep "loop A"; done
