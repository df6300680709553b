#!/bin/bash
# round 2, call BC (1 GPU): push with one ballot per queue instead of match.any
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export JIT=2 WAVES=4 MODES=1 FFR_JIT_NO_DISK_CACHE=1
{ for rep in 1 2; do for a in 1 0; do echo "== BALLOT_PEERS $a"; FFR_JIT_BALLOT_PEERS=$a python tools/probe.py csci tkoz3 2>&1 | cut -c1-100; done; done; } | tee gpurun_out/r2bc_probe.log
