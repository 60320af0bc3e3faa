#!/bin/bash
# ncu evidence for profiles/: one whole step (launch list) and --set full of the three dominant kernels.
# usage: gpurun --timeout 1900 -- 'bash tools/gpu_profile.sh <tag>'
tag=${1:-vX}
timeout 300 python bench.py --profile-step --no-cpu-baseline > gpurun_out/ps.json 2> gpurun_out/ps.err && \
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"k_mmp|k_dp_fill|k_dp_tb" -c 3 -o gpurun_out/r01_${tag}_top3 python bench.py --profile-step --no-cpu-baseline > gpurun_out/ncu_top3.log 2>&1
tail -2 gpurun_out/ncu_top3.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_${tag}.csv python bench.py --profile-step --no-cpu-baseline > gpurun_out/ncu_ps.log 2>&1
