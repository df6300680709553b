#!/bin/bash
# round 2, call C: K1d stats diet -- parity subset, thread-count / FMA probes, one ncu capture
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_jit.py tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -3 gpurun_out/r2c_pytest.log
export WAVES=4 JIT=2
for tpb in 320 352 384; do
  for fm in 0 1; do
    echo "== tpb $tpb fmad $fm" >> gpurun_out/r2c_probe.log
    FFR_JIT_TPB=$tpb FFR_JIT_NS=512 FFR_JIT_FMAD=$fm timeout 300 python tools/probe.py csci tkoz3 >> gpurun_out/r2c_probe.log 2>&1
  done
done
echo "== default (auto tpb)" >> gpurun_out/r2c_probe.log
timeout 300 python tools/probe.py csci tkoz3 csci8k >> gpurun_out/r2c_probe.log 2>&1
cat gpurun_out/r2c_probe.log
export FFR_JIT_DUMP_DIR=/tmp/ffrjit FFR_JIT_NO_DISK_CACHE=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ffr_jit_render -s 1 -c 1 -f -o gpurun_out/r2c_k1d_csci python tools/prof_one.py csci 0 2 2048 2 > gpurun_out/r2c_ncu.log 2>&1; tail -1 gpurun_out/r2c_ncu.log
