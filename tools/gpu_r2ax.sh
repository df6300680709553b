#!/bin/bash
# round 2, call AX (1 GPU): handkerchief / ex origin branch out of line (1.4 KB less code in the xform bodies)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export JIT=2 WAVES=4 MODES=1 FFR_JIT_NO_DISK_CACHE=1
{ for rep in 1 2; do for a in 1 0; do echo "== ORIGIN_OOL $a"; FFR_JIT_ORIGIN_OOL=$a python tools/probe.py csci tkoz3 2>&1 | cut -c1-100; done; done
echo "== ORIGIN_OOL 1 GEN 1"; FFR_JIT_ORIGIN_OOL=1 FFR_JIT_GEN_ROLLED=1 python tools/probe.py csci 2>&1 | cut -c1-100
echo "== ORIGIN_OOL 1 GEN 2"; FFR_JIT_ORIGIN_OOL=1 FFR_JIT_GEN_ROLLED=2 python tools/probe.py tkoz3 2>&1 | cut -c1-100; } | tee gpurun_out/r2ax_probe.log
