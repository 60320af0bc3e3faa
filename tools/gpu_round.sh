#!/bin/bash
# end-of-round sequence: all GPU tests, smoke, the reference arm, bench lines of every config with CPU legs, ncu evidence of cfg2
# usage: gpurun --timeout 5400 -- 'bash tools/gpu_round.sh <tag>'
tag=${1:-vX}
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
[ ${PIPESTATUS[0]} -eq 0 ] || exit 1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${tag}.json 2> gpurun_out/bench_ref.err; tail -c 400 gpurun_out/bench_ref_${tag}.json; echo
MP_BENCH_VERBOSE=1 timeout 900 python bench.py > gpurun_out/bench_cfg2_${tag}.json 2> gpurun_out/bench_cfg2.err; grep -E "timing|Elapsed time on host|loop R" gpurun_out/bench_cfg2.err | tail -14
for c in 1; do MP_BENCH_VERBOSE=1 timeout 400 python bench.py --no-cpu-baseline --steps 6 --contexts $c > gpurun_out/bench_cfg2_${tag}_contexts$c.json 2> gpurun_out/bq_ctx$c.err; grep "loop R" gpurun_out/bq_ctx$c.err; done
for cfg in cfg1 cfg4 cfg5; do MP_BENCH_VERBOSE=1 timeout 900 python bench.py --config $cfg --cli-pairs 0 --cpu-sample-pairs 100000 > gpurun_out/bench_${cfg}_${tag}.json 2> gpurun_out/bench_$cfg.err; grep "loop R" gpurun_out/bench_$cfg.err; done
MP_BENCH_VERBOSE=1 timeout 1800 python bench.py --config cfg3 --cli-pairs 2097152 --cpu-sample-pairs 100000 > gpurun_out/bench_cfg3_${tag}.json 2> gpurun_out/bench_cfg3.err; grep "loop R" gpurun_out/bench_cfg3.err
MP_BLOOM=0 MP_BENCH_VERBOSE=1 timeout 900 python bench.py --config cfg3 --no-cpu-baseline --steps 4 > gpurun_out/bq_cfg3_nobloom.json 2> gpurun_out/bq_cfg3_nobloom.err; echo "cfg3 without the K-mer filter: $(grep 'loop R' gpurun_out/bq_cfg3_nobloom.err)"
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/bench_*_${tag}*.json')):
    try:
        d=json.load(open(f))
        print(f, {k:d.get(k) for k in ('value','ms_per_step','parity_at_scale')}, d.get('e2e',{}).get('value'), (d.get('roofline') or {}).get('compute',{}).get('gcups_fill'), d.get('cpu_baseline',{}).get('value'), (d.get('e2e_cli') or {}).get('value'))
    except Exception as e: print(f, 'ERR', e)
PY
timeout 300 python bench.py --profile-step --no-cpu-baseline > gpurun_out/ps.json 2> gpurun_out/ps.err && \
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"k_mmp|k_dp_fill|k_dp_tb" -c 3 -o gpurun_out/r02_${tag}_top3 python bench.py --profile-step --no-cpu-baseline > gpurun_out/ncu_top3.log 2>&1
tail -2 gpurun_out/ncu_top3.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_${tag}.csv python bench.py --profile-step --no-cpu-baseline > gpurun_out/ncu_ps.log 2>&1
tail -1 gpurun_out/ncu_ps.log
timeout 600 python tools/driver_throughput.py 16 2>&1 | tail -8 | cut -c1-400 > gpurun_out/driver_throughput_${tag}.txt; cat gpurun_out/driver_throughput_${tag}.txt
MP_THR_NCU=1 timeout 600 python tools/driver_throughput.py 1 2>&1 | tail -3 | cut -c1-300
