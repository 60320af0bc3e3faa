#!/bin/bash
# quick loop on a GPU box: kernel-level parity tests + a short verbose bench (with and without the exact-occurrence shortcut)
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -5
MP_BENCH_VERBOSE=1 timeout 400 python bench.py --no-cpu-baseline --steps 9 > gpurun_out/bq.json 2> gpurun_out/bq.err; grep "loop R" gpurun_out/bq.err
MP_DP_EXACT=0 MP_BENCH_VERBOSE=1 timeout 400 python bench.py --no-cpu-baseline --steps 6 > gpurun_out/bq_noexact.json 2> gpurun_out/bq_noexact.err; grep "loop R" gpurun_out/bq_noexact.err
python - <<PY
import json
for f in ('gpurun_out/bq.json','gpurun_out/bq_noexact.json'):
    d=json.load(open(f))
    print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['compute']['gcups_fill'])
PY
