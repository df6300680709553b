#!/bin/bash
# round 2, call AR (1 GPU): stability of the K1d variants across process launches (same binary, 5 launches each)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export JIT=2 WAVES=4 MODES=1
for v in "0 1" "1 2" "0 2" "1 1"; do set -- $v
  for rep in 1 2 3 4 5; do echo "== ANG $1 GEN $2 rep $rep"; FFR_JIT_POLAR_ANG=$1 FFR_JIT_GEN_ROLLED=$2 python tools/probe.py csci tkoz3 2>&1 | cut -c1-100; done
done | tee gpurun_out/r2ar_probe.log | grep -v "^==" | awk '{print $1, $12}' | paste - - - - - - - - - - | head -8
