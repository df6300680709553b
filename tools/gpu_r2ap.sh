#!/bin/bash
# round 2, call AP (1 GPU): K1d with gen() fully rolled (one ISAAC step in the loop body) on top of polar per mask + sin/cos through sincos
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export JIT=2 WAVES=4 MODES=1 FFR_JIT_NO_DISK_CACHE=1 FFR_JIT_POLAR_NEED=1 FFR_JIT_SIN_VIA_SINCOS=1
for rep in 1 2; do for g in 1 2; do echo "== GEN $g"; FFR_JIT_GEN_ROLLED=$g python tools/probe.py csci tkoz3 2>&1 | cut -c1-120; done; done | tee gpurun_out/r2ap_probe.log
FFR_JIT_GEN_ROLLED=2 timeout 600 python -m pytest tests/test_gpu_jit.py -m gpu -q -x 2>&1 | tail -2
