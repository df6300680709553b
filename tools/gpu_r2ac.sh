#!/bin/bash
# round 2, call AC (1 GPU): which of the two layout changes costs what (K1d, csci6360 at 4096^2)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export JIT=2 WAVES=4 MODES=1 FFR_JIT_NO_DISK_CACHE=1
for v in "320 0 0" "320 1 0" "320 0 1" "320 1 1" "384 1 1" "352 1 1" "448 1 1" "384 1 0"; do
  set -- $v
  echo "== tpb $1 abc_global $2 stats_smem $3"
  FFR_JIT_TPB=$1 FFR_JIT_ABC_GLOBAL=$2 FFR_JIT_STATS_SMEM=$3 python tools/probe.py csci 2>&1 | cut -c1-200
  FFR_JIT_TPB=$1 FFR_JIT_ABC_GLOBAL=$2 FFR_JIT_STATS_SMEM=$3 python tools/probe.py csci 2>&1 | cut -c1-110
done 2>&1 | tee gpurun_out/r2ac_probe.log
