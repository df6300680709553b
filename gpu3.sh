JIT=2 timeout 300 python tools/probe.py csci tkoz3 2>&1 | grep "samples/s\|Error"
timeout 600 python -m pytest tests/test_gpu_jit.py -m gpu -q -x 2>&1 | tail -5
FFR_JIT_TPB=256 JIT=2 timeout 300 python tools/probe.py csci 2>&1 | grep "samples/s\|Error"
FFR_JIT_TPB=384 JIT=2 timeout 300 python tools/probe.py csci 2>&1 | grep "samples/s\|Error"
