#!/bin/bash
# round 2, call BE (1 GPU): randa/b/c in the L2 scratch -> more slots per block (576, 608) under the final code
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export JIT=2 WAVES=4 MODES=1 FFR_JIT_NO_DISK_CACHE=1
{ echo "== ABC 0"; python tools/probe.py csci tkoz3 2>&1 | cut -c1-170
echo "== ABC 1 (576)"; FFR_JIT_ABC_GLOBAL=1 python tools/probe.py csci tkoz3 2>&1 | cut -c1-170
echo "== ABC 1 NS 608"; FFR_JIT_ABC_GLOBAL=1 FFR_JIT_NS=608 python tools/probe.py csci tkoz3 2>&1 | cut -c1-170
echo "== ABC 1 NS 544"; FFR_JIT_ABC_GLOBAL=1 FFR_JIT_NS=544 python tools/probe.py csci tkoz3 2>&1 | cut -c1-170
echo "== ABC 1 NS 512"; FFR_JIT_ABC_GLOBAL=1 FFR_JIT_NS=512 python tools/probe.py csci tkoz3 2>&1 | cut -c1-170; } | tee gpurun_out/r2be_probe.log
