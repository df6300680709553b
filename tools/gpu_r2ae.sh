#!/bin/bash
# round 2, call AE (1 GPU): K1d with chain start, bad-value record and the idle path out of line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export JIT=2 WAVES=4 MODES=1 FFR_JIT_NO_DISK_CACHE=1
for i in 1 2; do python tools/probe.py csci tkoz3 2>&1 | cut -c1-150; done | tee gpurun_out/r2ae_probe.log
unset FFR_JIT_NO_DISK_CACHE
( timeout 900 python -m pytest tests/test_gpu_jit.py tests/test_gpu_f32.py -m gpu -q -x ) > gpurun_out/r2ae_pytest.log 2>&1
tail -3 gpurun_out/r2ae_pytest.log | cut -c1-200
