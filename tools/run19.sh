timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "dp_batch" --tb=short 2>&1 | tail -40 > gpurun_out/t19.log; tail -40 gpurun_out/t19.log | cut -c1-600
