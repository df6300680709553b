timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
