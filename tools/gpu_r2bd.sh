#!/bin/bash
# round 2, call BD (1 GPU): slots per block around the default (queue depth vs nothing else)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export JIT=2 WAVES=4 MODES=1 FFR_JIT_NO_DISK_CACHE=1
{ for n in 0 448 480 512 544; do echo "== NS $n"; FFR_JIT_NS=$n python tools/probe.py csci tkoz3 2>&1 | cut -c1-160; done; } | tee gpurun_out/r2bd_probe.log
