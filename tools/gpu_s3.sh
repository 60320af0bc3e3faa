#!/bin/bash
# all GPU tests, a quick cfg2 bench and the cfg4 bench line (DP-heavy: stages S2 / S3 matter) with its CPU leg + parity at scale
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
MP_BENCH_VERBOSE=1 timeout 400 python bench.py --no-cpu-baseline --steps 9 > gpurun_out/bq.json 2> gpurun_out/bq.err; grep "loop R" gpurun_out/bq.err
MP_TRACE=2 MP_BENCH_VERBOSE=1 timeout 900 python bench.py --config cfg4 --cli-pairs 0 --cpu-sample-pairs 100000 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
grep "mp_trace" gpurun_out/bench_cfg4.err | tail -24; grep "loop R" gpurun_out/bench_cfg4.err
python - <<PY
import json
for f in ('bq','bench_cfg4'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f, {k:d.get(k) for k in ('value','ms_per_step','parity_at_scale')}, d['e2e']['value'], d['roofline']['compute']['gcups_fill'], d.get('cpu_baseline',{}).get('value'), d['roofline']['stage_ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
