#!/bin/bash
# DP-focused loop: kernel-level parity tests FIRST (nothing else runs if they fail), then short benches of cfg2 (with and without the
# exact-occurrence shortcut), cfg4, cfg5
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -5
[ ${PIPESTATUS[0]} -eq 0 ] || exit 1
for cfg in cfg2 cfg4 cfg5; do MP_BENCH_VERBOSE=1 timeout 300 python bench.py --config $cfg --no-cpu-baseline --steps 6 > gpurun_out/bq_$cfg.json 2> gpurun_out/bq_$cfg.err; grep "loop R" gpurun_out/bq_$cfg.err; done
MP_DP_EXACT=0 MP_BENCH_VERBOSE=1 timeout 300 python bench.py --no-cpu-baseline --steps 6 > gpurun_out/bq_noexact.json 2> gpurun_out/bq_noexact.err; grep "loop R" gpurun_out/bq_noexact.err
python - <<PY
import json
for f in ('bq_cfg2','bq_cfg4','bq_cfg5','bq_noexact'):
    d=json.load(open('gpurun_out/%s.json'%f))
    print(f, {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['compute']['gcups_fill'])
PY
