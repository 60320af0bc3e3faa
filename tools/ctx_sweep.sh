#!/bin/bash
# contexts-per-GPU sweep of the bench (value / e2e only)
for c in 3 4; do
timeout 300 python bench.py --no-cpu-baseline --steps 12 --contexts $c > gpurun_out/bq_c$c.json 2> gpurun_out/bq_c$c.err; python - <<PY
import json
d=json.load(open('gpurun_out/bq_c$c.json'))
print("contexts $c", round(d['value']/1e6,2), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,2))
PY
done
