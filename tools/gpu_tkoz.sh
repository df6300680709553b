mkdir -p gpurun_out
export JIT=2 WAVES=2
MODES=1,4 python tools/probe.py tkoz3 csci 2>&1 | cut -c1-110 > gpurun_out/tkoz_probe.log
cat gpurun_out/tkoz_probe.log
