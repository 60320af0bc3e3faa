export MP_BENCH_VERBOSE=1 MP_BENCH_CONTEXTS=1
for mode in "MP_BENCH_NO_SAMPLER=1" "MP_BENCH_SAMPLE_MS=200" "MP_BENCH_SAMPLE_MS=1000"; do
  echo "== $mode"; env $mode python bench.py --no-cpu-baseline --steps 8 2>&1 >/dev/null | grep "ctx " | awk '{printf "%s ", $7}'; echo
done
