timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
REGROUP=2 timeout 600 python tools/probe.py csci tkoz3 2>&1 | tail -5
