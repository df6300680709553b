#!/bin/bash
# round 2, call AZ (1 GPU): sin/cos of r and r^2 as shared polar quantities (one two-argument sincos call where an xform needs both)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export JIT=2 WAVES=4 MODES=1 FFR_JIT_NO_DISK_CACHE=1
{ for rep in 1 2; do for a in 1 0; do echo "== SINR $a"; FFR_JIT_SINR=$a python tools/probe.py csci tkoz3 2>&1 | cut -c1-100; done; done
echo "== SINR 1 GEN 1"; FFR_JIT_SINR=1 FFR_JIT_GEN_ROLLED=1 python tools/probe.py csci 2>&1 | cut -c1-100; } | tee gpurun_out/r2az_probe.log
unset FFR_JIT_NO_DISK_CACHE
( FFR_JIT_SINR=1 timeout 900 python -m pytest tests/test_gpu_jit.py -m gpu -q -x ) 2>&1 | tail -2
