#!/bin/bash
# round 2, call AV (1 GPU): K1d with ONE table-driven copy of the pre and post affines in the xform dispatch
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export JIT=2 WAVES=4 MODES=1 FFR_JIT_NO_DISK_CACHE=1
{ for rep in 1 2; do for a in 1 0; do echo "== AFFTAB $a"; FFR_JIT_AFFINE_TAB=$a python tools/probe.py csci tkoz3 2>&1 | cut -c1-100; done; done
echo "== AFFTAB 1 GEN 1"; FFR_JIT_AFFINE_TAB=1 FFR_JIT_GEN_ROLLED=1 python tools/probe.py csci 2>&1 | cut -c1-100
echo "== AFFTAB 1 GEN 2"; FFR_JIT_AFFINE_TAB=1 FFR_JIT_GEN_ROLLED=2 python tools/probe.py tkoz3 2>&1 | cut -c1-100; } | tee gpurun_out/r2av_probe.log
unset FFR_JIT_NO_DISK_CACHE
( timeout 900 python -m pytest tests/test_gpu_jit.py -m gpu -q -x ) 2>&1 | tail -2
