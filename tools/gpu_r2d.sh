#!/bin/bash
# round 2, call D: K1d block-shape sweep (one big block per SM vs two), after the MODES split
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_jit.py tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -3 gpurun_out/r2d_pytest.log
export WAVES=4 JIT=2
for cfg in "320 2 0" "320 2 0" "640 1 0" "704 1 0" "768 1 0" "512 1 768" "576 1 832" "640 1 896" "288 2 512" "256 3 0"; do
  set -- $cfg
  echo "== tpb $1 minb $2 ns $3" >> gpurun_out/r2d_probe.log
  FFR_JIT_TPB=$1 FFR_JIT_MINB=$2 FFR_JIT_NS=$3 timeout 300 python tools/probe.py csci tkoz3 >> gpurun_out/r2d_probe.log 2>&1
done
cat gpurun_out/r2d_probe.log
