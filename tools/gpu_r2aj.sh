#!/bin/bash
# round 2, call AJ (1 GPU): K1d with the polar unit specialised per need mask and one shared reciprocal for its two divisions
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export JIT=2 WAVES=4 MODES=1 FFR_JIT_NO_DISK_CACHE=1
for i in 1 2; do python tools/probe.py csci tkoz3 2>&1 | cut -c1-150; done | tee gpurun_out/r2aj_probe.log
FFR_JIT_POLAR_NEED=0 python tools/probe.py csci tkoz3 2>&1 | cut -c1-150 | tee -a gpurun_out/r2aj_probe.log
unset FFR_JIT_NO_DISK_CACHE
( timeout 1200 python -m pytest tests/test_gpu_jit.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x ) > gpurun_out/r2aj_pytest.log 2>&1
tail -3 gpurun_out/r2aj_pytest.log | cut -c1-200
