#!/bin/bash
# bench line of cfg3 (8 Gbp multi-genome text, soap4-nt2.ini), CPU legs included
MP_TRACE=2 MP_BENCH_VERBOSE=1 timeout 2400 python bench.py --config cfg3 --cli-pairs 2097152 --cpu-sample-pairs 100000 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -5 gpurun_out/bench_cfg3.err
grep "mp_trace" gpurun_out/bench_cfg3.err | tail -30; grep "loop R" gpurun_out/bench_cfg3.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_cfg3.json'))
print({k:d.get(k) for k in ('value','ms_per_step','parity_at_scale','aligned_fraction','reads_with_score_ge_40_fraction','index_prepare_s')}, d['e2e']['value'], d.get('cpu_baseline'), d.get('e2e_cli',{}).get('value'), d['roofline']['stage_ms_per_step'], d['config']['l2'])
print(d['roofline']['seeding'])
PY
