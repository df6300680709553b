#!/bin/bash
# round 2, call H: K1d lane box (merged at kernel end) and exact affine rows, A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export WAVES=4 JIT=2
for f in 0 1 0 1; do
  echo "== affine full $f" >> gpurun_out/r2h_probe.log
  FFR_JIT_AFFINE_FULL=$f timeout 300 python tools/probe.py csci tkoz3 >> gpurun_out/r2h_probe.log 2>&1
done
cat gpurun_out/r2h_probe.log
( timeout 1200 python -m pytest tests/test_gpu_jit.py tests/test_gpu_parity.py -m gpu -x -q ) > gpurun_out/r2h_pytest.log 2>&1
tail -3 gpurun_out/r2h_pytest.log
