JIT=2 timeout 300 python tools/probe.py tkoz3 csci 2>&1 | grep "samples/s\|Error"
