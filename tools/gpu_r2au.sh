#!/bin/bash
# round 2, call AU (1 GPU): the float build under the new K1d defaults (gen form x thread count)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export JIT=2 WAVES=4 MODES=1 ELEM=4 FFR_JIT_NO_DISK_CACHE=1
{ echo "== default"; python tools/probe.py csci tkoz3 2>&1 | cut -c1-150
echo "== GEN 1"; FFR_JIT_GEN_ROLLED=1 python tools/probe.py csci tkoz3 2>&1 | cut -c1-150
echo "== GEN 2"; FFR_JIT_GEN_ROLLED=2 python tools/probe.py csci tkoz3 2>&1 | cut -c1-150
echo "== ANG 0"; FFR_JIT_POLAR_ANG=0 python tools/probe.py csci tkoz3 2>&1 | cut -c1-150
echo "== POLAR_NEED 0"; FFR_JIT_POLAR_NEED=0 python tools/probe.py csci tkoz3 2>&1 | cut -c1-150; } | tee gpurun_out/r2au_probe.log
