#!/bin/bash
# round 2, call S: K1d with chain start and bad-value handling out of line (hot code contiguous)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export WAVES=4 JIT=2
timeout 300 python tools/probe.py csci tkoz3 > gpurun_out/r2s_probe.log 2>&1

timeout 300 python tools/probe.py csci tkoz3 >> gpurun_out/r2s_probe.log 2>&1
cat gpurun_out/r2s_probe.log
unset WAVES JIT
( timeout 900 python -m pytest tests/test_gpu_jit.py tests/test_gpu_parity.py -m gpu -x -q ) > gpurun_out/r2s_pytest.log 2>&1
tail -3 gpurun_out/r2s_pytest.log
