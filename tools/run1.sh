set -x
nvidia-smi -L; nproc; free -g | head -2; df -h /tmp | tail -1
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --ref-mbp 50 --pairs-per-step 262144 --steps 2 --warmup 1 --cpu-sample-pairs 50000 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo rc=$?; tail -3 gpurun_out/bench_small.err; cat gpurun_out/bench_small.json
