export MP_BENCH_VERBOSE=1
for c in 1 2; do echo "== contexts $c"; MP_BENCH_CONTEXTS=$c python bench.py --no-cpu-baseline --steps 8 > gpurun_out/bc$c.json 2> gpurun_out/bc$c.err; grep "ctx " gpurun_out/bc$c.err | awk '{printf "%s ", $7}'; echo; python - <<PY
import json
d=json.load(open('gpurun_out/bc$c.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])
PY
done
