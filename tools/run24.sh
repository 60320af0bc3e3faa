timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/big_e2e.py 2400000 2>&1 | tail -5
