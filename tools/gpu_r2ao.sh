#!/bin/bash
# round 2, call AO (1 GPU): K1d code-size search: polar per mask + sin/cos through sincos fixed on,
# cold paths out of line in all 8 combinations (bit 0 chain start, 1 idle path, 2 bad-value record)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export JIT=2 WAVES=4 MODES=1 FFR_JIT_NO_DISK_CACHE=1 FFR_JIT_POLAR_NEED=1 FFR_JIT_SIN_VIA_SINCOS=1
for c in 0 1 2 3 4 5 6 7; do echo "== COLD $c"; FFR_JIT_COLD=$c python tools/probe.py csci tkoz3 2>&1 | cut -c1-120; done | tee gpurun_out/r2ao_probe.log
