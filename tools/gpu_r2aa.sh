#!/bin/bash
# round 2, call AA (1 GPU): K1d with six warps per scheduler (384 threads x 2 blocks, 80 registers,
# a/b/c in the L2 scratch, per-warp statistics in shared memory): thread-count sweep + JIT tests
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export JIT=2 WAVES=2 MODES=1 FFR_JIT_NO_DISK_CACHE=1
{ for t in 384 352 320; do echo "== TPB $t"; FFR_JIT_TPB=$t python tools/probe.py csci tkoz3 csci8k; done; } 2>&1 | cut -c1-230 > gpurun_out/r2aa_probe.log; cat gpurun_out/r2aa_probe.log
unset FFR_JIT_NO_DISK_CACHE
( timeout 900 python -m pytest tests/test_gpu_jit.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_gpu_f32.py -m gpu -q -x ) > gpurun_out/r2aa_pytest.log 2>&1
tail -4 gpurun_out/r2aa_pytest.log | cut -c1-200
