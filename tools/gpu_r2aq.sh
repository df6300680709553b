#!/bin/bash
# round 2, call AQ (1 GPU): atan2 called from the xform body (polar copies shared across NEED_ANG) vs from inside the polar unit
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export JIT=2 WAVES=4 MODES=1 FFR_JIT_NO_DISK_CACHE=1
for rep in 1 2; do for a in 1 0; do echo "== ANG $a"; FFR_JIT_POLAR_ANG=$a python tools/probe.py csci tkoz3 2>&1 | cut -c1-120; done; done | tee gpurun_out/r2aq_probe.log
for a in 1; do echo "== ANG $a GEN 2"; FFR_JIT_GEN_ROLLED=2 FFR_JIT_POLAR_ANG=$a python tools/probe.py csci tkoz3 2>&1 | cut -c1-120; done | tee -a gpurun_out/r2aq_probe.log
