set -x
timeout 300 python bench.py --profile-step --no-cpu-baseline > gpurun_out/ps.json 2> gpurun_out/ps.err && \
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step.csv python bench.py --profile-step --no-cpu-baseline > gpurun_out/ncu_ps.log 2>&1
tail -2 gpurun_out/ncu_ps.log; wc -l gpurun_out/launches_step.csv
