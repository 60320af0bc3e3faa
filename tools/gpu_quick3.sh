#!/bin/bash
# kernel-level parity tests, a short cfg2 bench and a short cfg3 bench (no CPU legs)
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3
[ ${PIPESTATUS[0]} -eq 0 ] || exit 1
MP_BENCH_VERBOSE=1 timeout 400 python bench.py --no-cpu-baseline --steps 9 > gpurun_out/bq.json 2> gpurun_out/bq.err; grep "loop R" gpurun_out/bq.err
MP_BENCH_VERBOSE=1 timeout 1200 python bench.py --config cfg3 --no-cpu-baseline --steps 6 > gpurun_out/bq_cfg3.json 2> gpurun_out/bq_cfg3.err; grep "loop R" gpurun_out/bq_cfg3.err
python - <<PY
import json
for f in ('bq','bq_cfg3'):
    d=json.load(open('gpurun_out/%s.json'%f))
    print(f, {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['stage_ms_per_step'])
PY
