timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
REGROUP=1,2 timeout 600 python tools/probe.py tkoz3 csci barnsley 2>&1 | tail -20
