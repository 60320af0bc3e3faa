#!/bin/bash
# all GPU tests + a verbose bench
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
MP_BENCH_VERBOSE=1 timeout 400 python bench.py --no-cpu-baseline --steps 9 > gpurun_out/bq.json 2> gpurun_out/bq.err; grep "loop R" gpurun_out/bq.err
python - <<PY
import json
d=json.load(open('gpurun_out/bq.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])
PY
