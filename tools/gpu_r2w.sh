#!/bin/bash
# round 2, call W: constant-bank atan2: accuracy + parity tests, probe
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_jit.py tests/test_gpu_f32.py -m gpu -x -q ) > gpurun_out/r2w_pytest.log 2>&1
tail -5 gpurun_out/r2w_pytest.log
export WAVES=4 JIT=2
timeout 300 python tools/probe.py csci tkoz3 > gpurun_out/r2w_probe.log 2>&1
timeout 300 python tools/probe.py csci tkoz3 >> gpurun_out/r2w_probe.log 2>&1
cat gpurun_out/r2w_probe.log
