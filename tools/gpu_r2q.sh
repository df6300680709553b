#!/bin/bash
# round 2, call Q: ncu of K1e + compact tile with block rows on sierpinski_3d@512^3
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export FFR_JIT_DUMP_DIR=/tmp/ffrjit FFR_JIT_NO_DISK_CACHE=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ffr_jit_render -s 1 -c 1 -f -o gpurun_out/r2q_k1e_s3d python tools/prof_one.py sierp3d 0 2 2048 2 > gpurun_out/r2q_ncu.log 2>&1; tail -1 gpurun_out/r2q_ncu.log
