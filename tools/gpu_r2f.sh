#!/bin/bash
# round 2, call F: K1d queue rewrite (lap parity, release/acquire, predicated atomics): tests, A/B, ncu
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_jit.py -m gpu -x -q --durations=5 ) > gpurun_out/r2f_pytest.log 2>&1
tail -9 gpurun_out/r2f_pytest.log
export WAVES=4 JIT=2
for a in 1 0 1 0; do
  echo "== acqrel $a" >> gpurun_out/r2f_probe.log
  FFR_JIT_ACQREL=$a timeout 300 python tools/probe.py csci tkoz3 >> gpurun_out/r2f_probe.log 2>&1
done
cat gpurun_out/r2f_probe.log
export FFR_JIT_DUMP_DIR=/tmp/ffrjit FFR_JIT_NO_DISK_CACHE=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ffr_jit_render -s 1 -c 1 -f -o gpurun_out/r2f_k1d_csci python tools/prof_one.py csci 0 2 2048 2 > gpurun_out/r2f_ncu.log 2>&1; tail -1 gpurun_out/r2f_ncu.log
