#!/bin/bash
# round 2, call BA (1 GPU): the new division soak test (default size, then 2e11 samples)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_jit.py -m gpu -q -x --tb=short -k "shared_reciprocal" ) > gpurun_out/r2ba_div.log 2>&1
tail -30 gpurun_out/r2ba_div.log | cut -c1-220
( time FFR_DIV_SOAK_SAMPLES=2e11 timeout 900 python -m pytest tests/test_gpu_jit.py -m gpu -q -k "shared_reciprocal and 17" ) 2>&1 | tail -6 | tee -a gpurun_out/r2ba_div.log
