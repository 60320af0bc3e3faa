#!/bin/bash
# The round-end sequence on a GPU box, in the driver's order: GPU parity tests, smoke(), the reference arm (fresh box: it has to
# prepare the workload through its child process), then the default bench.
# usage (from the repo root): gpurun --timeout 2700 -- 'bash tools/gpu_check.sh'
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
