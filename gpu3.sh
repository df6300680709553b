timeout 900 python -m pytest tests/test_gpu_jit.py -m gpu -q -x 2>&1 | tail -15
JIT=1,2 timeout 600 python tools/probe.py csci tkoz3 barnsley 2>&1 | tail -20
