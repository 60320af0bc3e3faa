#!/bin/bash
# all GPU tests + benches of several configs (no CPU legs except cfg2)
set -x
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for c in 1 3; do MP_BENCH_VERBOSE=1 timeout 400 python bench.py --no-cpu-baseline --steps 6 --contexts $c > gpurun_out/bq_ctx$c.json 2> gpurun_out/bq_ctx$c.err; grep "loop R" gpurun_out/bq_ctx$c.err; done
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; grep -E "timing|Elapsed time on host" gpurun_out/bench.err | tail -12
for cfg in cfg1 cfg4 cfg5; do timeout 900 python bench.py --config $cfg --cli-pairs 0 --cpu-sample-pairs 100000 > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err; tail -2 gpurun_out/bench_$cfg.err; done
python - <<PY
import json
for f in ('bq_ctx1','bq_ctx3','bench','bench_cfg1','bench_cfg4','bench_cfg5'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f, {k:d.get(k) for k in ('value','ms_per_step','parity_at_scale')}, d['e2e']['value'], d['roofline']['compute']['gcups_fill'], d.get('cpu_baseline',{}).get('value'), d.get('e2e_cli',{}).get('value'), d['roofline']['stage_ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
