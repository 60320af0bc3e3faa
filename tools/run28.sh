for c in 2 3; do
timeout 400 python bench.py --no-cpu-baseline --steps 9 --contexts $c > gpurun_out/bq_c$c.json 2> gpurun_out/bq_c$c.err; python - <<PY
import json
d=json.load(open('gpurun_out/bq_c$c.json'))
print("contexts $c", d['value'], d['ms_per_step'], d['e2e']['value'])
PY
done
