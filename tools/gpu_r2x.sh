#!/bin/bash
# round 2, call X: K1d block-shape sweep on the final build (gen rolled, queue rewrite), atan2 accuracy test
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k atan2 ) > gpurun_out/r2x_pytest.log 2>&1
tail -3 gpurun_out/r2x_pytest.log
export WAVES=4 JIT=2
for cfg in "320 2 0" "352 2 512" "384 2 512" "288 2 512" "640 1 0" "768 1 0" "256 3 0" "320 2 480" "320 2 448"; do
  set -- $cfg
  echo "== tpb $1 minb $2 ns $3" >> gpurun_out/r2x_probe.log
  FFR_JIT_TPB=$1 FFR_JIT_MINB=$2 FFR_JIT_NS=$3 timeout 300 python tools/probe.py csci tkoz3 >> gpurun_out/r2x_probe.log 2>&1
done
cut -c1-175 gpurun_out/r2x_probe.log
