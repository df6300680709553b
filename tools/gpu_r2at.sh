#!/bin/bash
# round 2, call AT (1 GPU): sincos unit with the far path behind a call (smaller unit) vs inline
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export JIT=2 WAVES=4 MODES=1 FFR_JIT_NO_DISK_CACHE=1
{ for rep in 1 2; do for f in 0 1; do echo "== FARCALL $f"; FFR_JIT_SINCOS_FARCALL=$f python tools/probe.py csci tkoz3 2>&1 | cut -c1-100; done; done
echo "== FARCALL 1 GEN 1"; FFR_JIT_SINCOS_FARCALL=1 FFR_JIT_GEN_ROLLED=1 python tools/probe.py csci 2>&1 | cut -c1-100
echo "== FARCALL 1 GEN 2 tkoz"; FFR_JIT_SINCOS_FARCALL=1 FFR_JIT_GEN_ROLLED=2 python tools/probe.py tkoz3 2>&1 | cut -c1-100; } | tee gpurun_out/r2at_probe.log
