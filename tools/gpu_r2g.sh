#!/bin/bash
# round 2, call G: K1d lane box + exact affine rows: tests, probe
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_jit.py tests/test_gpu_parity.py tests/test_gpu_f32.py -m gpu -x -q ) > gpurun_out/r2g_pytest.log 2>&1
tail -4 gpurun_out/r2g_pytest.log
export WAVES=4 JIT=2
for a in 1 2; do
  timeout 300 python tools/probe.py csci tkoz3 csci8k >> gpurun_out/r2g_probe.log 2>&1
done
cat gpurun_out/r2g_probe.log
