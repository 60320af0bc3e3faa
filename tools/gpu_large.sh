timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "index_build" 2>&1 | tail -8
timeout 1500 python -m pytest tests/test_gpu_e2e.py -m gpu -x -q -k "beyond" --durations=3 2>&1 | tail -25
