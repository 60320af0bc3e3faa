#!/bin/bash
# GPU parity tests, smoke(), a default bench, then ncu evidence (launch list of one step + --set full of the top kernels).
# usage: gpurun --timeout 2700 -- 'bash tools/gpu_verify_profile.sh <tag>'
tag=${1:-vX}
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
MP_BENCH_VERBOSE=1 timeout 900 python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; tail -c 1200 gpurun_out/bench_${tag}.json; grep "loop R" gpurun_out/bench_${tag}.err
timeout 300 python bench.py --profile-step --no-cpu-baseline > gpurun_out/ps.json 2> gpurun_out/ps.err && \
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"k_mmp|k_dp_fill|k_dp_tb" -c 3 -o gpurun_out/r02_${tag}_top3 python bench.py --profile-step --no-cpu-baseline > gpurun_out/ncu_top3.log 2>&1
tail -2 gpurun_out/ncu_top3.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_${tag}.csv python bench.py --profile-step --no-cpu-baseline > gpurun_out/ncu_ps.log 2>&1
tail -2 gpurun_out/ncu_ps.log
