timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k sin_cos 2>&1 | grep -E "assert|Error|passed|failed|^E " | head -12
