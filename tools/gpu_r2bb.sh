#!/bin/bash
# round 2, call BB (1 GPU): the 1e12-sample queue soak on the final build
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time FFR_SOAK_SAMPLES=1e12 timeout 900 python -m pytest tests/test_gpu_jit.py -m gpu -q -k soak ) > gpurun_out/r2bb_soak.log 2>&1
tail -6 gpurun_out/r2bb_soak.log
