#!/bin/bash
# round 2, call BF (1 GPU): where a warp's queue scan starts (instruction-cache sharing between warps)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export JIT=2 WAVES=4 MODES=1 FFR_JIT_NO_DISK_CACHE=1
{ for m in 0 1 2; do echo "== ROT_MODE $m"; FFR_JIT_ROT_MODE=$m python tools/probe.py csci tkoz3 2>&1 | cut -c1-100; done; } | tee gpurun_out/r2bf_probe.log
