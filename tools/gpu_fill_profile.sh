#!/bin/bash
# ncu --set full of the DP fill kernel (one step of the bench workload) + a plain bench with the exact-occurrence shortcut off
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_fill_profile.sh <tag>'
tag=${1:-vX}
MP_DP_EXACT=0 MP_BENCH_VERBOSE=1 timeout 400 python bench.py --no-cpu-baseline --steps 6 > gpurun_out/bq_noexact_${tag}.json 2> gpurun_out/bq_noexact.err; grep "loop R" gpurun_out/bq_noexact.err
timeout 300 python bench.py --profile-step --no-cpu-baseline > gpurun_out/ps.json 2> gpurun_out/ps.err && \
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"k_dp_fill" -c 2 -o gpurun_out/r02_${tag}_fill python bench.py --profile-step --no-cpu-baseline > gpurun_out/ncu_fill.log 2>&1
tail -2 gpurun_out/ncu_fill.log
