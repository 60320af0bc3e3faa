#!/bin/bash
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for cfg in cfg3 cfg4; do MP_TRACE=2 MP_BENCH_VERBOSE=1 timeout 900 python bench.py --config $cfg --no-cpu-baseline --steps 4 > gpurun_out/bq_$cfg.json 2> gpurun_out/bq_$cfg.err; grep "mp_trace.*s2\|mp_trace.*s3\|mp_trace. seed_pairs\|mp_trace. deep" gpurun_out/bq_$cfg.err | tail -9; grep "loop R" gpurun_out/bq_$cfg.err; done
python - <<PY
import json
for f in ('bq_cfg3','bq_cfg4'):
    d=json.load(open('gpurun_out/%s.json'%f))
    print(f, {k:d.get(k) for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['stage_ms_per_step'])
PY
