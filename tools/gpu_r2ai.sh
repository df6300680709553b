#!/bin/bash
# round 2, call AI (1 GPU): float build with 448 threads by default: tests + probe
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_f32.py tests/test_gpu_jit.py -m gpu -q ) > gpurun_out/r2ai_pytest.log 2>&1
tail -3 gpurun_out/r2ai_pytest.log | cut -c1-200
export JIT=2 WAVES=4 MODES=1 ELEM=4 FFR_JIT_NO_DISK_CACHE=1
python tools/probe.py csci tkoz3 sierpinski barnsley sierp3d 2>&1 | cut -c1-230 | tee gpurun_out/r2ai_probe.log
