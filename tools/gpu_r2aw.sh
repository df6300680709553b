#!/bin/bash
# round 2, call AW (1 GPU): ncu --set full of K1d on the final build (csci6360, tkoz_test3 at 4096^2)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export FFR_JIT_DUMP_DIR=/tmp/ffrjit FFR_JIT_NO_DISK_CACHE=1
for cfg in csci tkoz3; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:ffr_jit_render -s 1 -c 1 -f -o gpurun_out/r2aw_$cfg python tools/prof_one.py $cfg 0 2 2048 2 > gpurun_out/r2aw_ncu_$cfg.log 2>&1; tail -1 gpurun_out/r2aw_ncu_$cfg.log
done
ls -la gpurun_out/r2aw_*.ncu-rep
