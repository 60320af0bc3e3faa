export MP_BENCH_VERBOSE=1
for c in 1 2; do MP_BENCH_CONTEXTS=$c python bench.py --no-cpu-baseline --steps 6 > gpurun_out/b$c.json 2> gpurun_out/b$c.err; grep "ctx " gpurun_out/b$c.err | awk '{print $2,$4,$5,$7}' | tr '\n' ';'; echo; python - <<PY
import json
d=json.load(open('gpurun_out/b$c.json'))
print("$c contexts:", {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])
PY
done
