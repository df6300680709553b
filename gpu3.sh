JIT=2 timeout 300 python tools/probe.py tkoz3 csci 2>&1 | grep "samples/s\|Error"
FFR_JIT_TPB=320 JIT=2 timeout 300 python tools/probe.py tkoz3 csci 2>&1 | grep "samples/s\|Error"
timeout 900 python -m pytest tests/test_gpu_jit.py -m gpu -q -x 2>&1 | tail -5
