#!/bin/bash
# The round-end sequence on a GPU box: GPU parity tests, smoke(), the default bench and the reference arm.
# usage (from the repo root): gpurun --timeout 2400 -- 'bash tools/gpu_check.sh'
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 300 gpurun_out/bench.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.json
