#!/bin/bash
# round 2, call AK (1 GPU): polar unit variants: 0 = as before, 1 = one copy per need mask + shared reciprocal, 2 = one copy + shared reciprocal
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export JIT=2 WAVES=4 MODES=1 FFR_JIT_NO_DISK_CACHE=1
for m in 0 2 1 0 2 1; do echo "== polar $m"; FFR_JIT_POLAR_NEED=$m python tools/probe.py csci tkoz3 csci8k 2>&1 | cut -c1-120; done | tee gpurun_out/r2ak_probe.log
