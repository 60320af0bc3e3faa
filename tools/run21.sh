timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err; tail -c 600 gpurun_out/bench_v5.json
