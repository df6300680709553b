mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_affine.py tests/test_gpu_jit.py -m gpu -x -q ) > gpurun_out/acc_pytest.log 2>&1; tail -5 gpurun_out/acc_pytest.log
export JIT=2 WAVES=4 MODES=1 FFR_JIT_NO_DISK_CACHE=1
for mb in 0 32 1024; do echo "== FFR_ACC_MAX_MB $mb"; FFR_ACC_MAX_MB=$mb python tools/probe.py sierpinski barnsley1k barnsley barnsley4k sierp4k sierp3d256 2>&1 | cut -c1-104; done > gpurun_out/acc_probe.log 2>&1
cat gpurun_out/acc_probe.log
