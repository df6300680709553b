#!/bin/bash
# round 2, call AS (1 GPU): the remaining switches around the new default point of K1d
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export JIT=2 WAVES=4 MODES=1 FFR_JIT_NO_DISK_CACHE=1
{ echo "== default"; python tools/probe.py csci tkoz3 2>&1 | cut -c1-100
for c in 1 2 3 4 5 6 7; do echo "== COLD $c"; FFR_JIT_COLD=$c python tools/probe.py csci tkoz3 2>&1 | cut -c1-100; done
echo "== ROT_STATIC 0"; FFR_JIT_ROT_STATIC=0 python tools/probe.py csci tkoz3 2>&1 | cut -c1-100
echo "== SC_XOR 0"; FFR_SC_XOR=0 python tools/probe.py csci tkoz3 2>&1 | cut -c1-100
echo "== TPB 288"; FFR_JIT_TPB=288 python tools/probe.py csci tkoz3 2>&1 | cut -c1-100
echo "== default"; python tools/probe.py csci tkoz3 2>&1 | cut -c1-100; } | tee gpurun_out/r2as_probe.log
