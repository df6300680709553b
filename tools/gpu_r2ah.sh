#!/bin/bash
# round 2, call AH (1 GPU): does the warp count help where registers and slots are not the limit? (float build)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export JIT=2 WAVES=4 MODES=1 ELEM=4 FFR_JIT_NO_DISK_CACHE=1
for t in 256 320 384 448 512; do echo "== TPB $t"; FFR_JIT_TPB=$t python tools/probe.py csci 2>&1 | cut -c1-200; done | tee gpurun_out/r2ah_probe.log
