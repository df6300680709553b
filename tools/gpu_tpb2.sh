mkdir -p gpurun_out
export JIT=2 WAVES=2 MODES=1 FFR_JIT_NO_DISK_CACHE=1
{ for t in 320 256 224 288; do echo "== TPB $t"; FFR_JIT_TPB=$t python tools/probe.py csci tkoz3; done; } 2>&1 | cut -c1-200 > gpurun_out/tpb2_probe.log; cat gpurun_out/tpb2_probe.log
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sin_cos or single or ieee" ) > gpurun_out/tpb2_pytest.log 2>&1; tail -2 gpurun_out/tpb2_pytest.log
