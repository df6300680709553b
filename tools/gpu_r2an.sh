#!/bin/bash
# round 2, call AN (1 GPU): code-size trade-offs on K1d: polar per mask (P) x sin/cos through sincos (S)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export JIT=2 WAVES=4 MODES=1 FFR_JIT_NO_DISK_CACHE=1
for rep in 1 2; do for v in "0 0" "1 0" "0 1" "1 1"; do set -- $v; echo "== P $1 S $2"; FFR_JIT_POLAR_NEED=$1 FFR_JIT_SIN_VIA_SINCOS=$2 python tools/probe.py csci tkoz3 2>&1 | cut -c1-120; done; done | tee gpurun_out/r2an_probe.log
unset FFR_JIT_NO_DISK_CACHE
( timeout 600 python -m pytest tests/test_cpp_mirror.py -m gpu -q ) > gpurun_out/r2an_pytest.log 2>&1; tail -3 gpurun_out/r2an_pytest.log | cut -c1-200
