mkdir -p gpurun_out
export JIT=2 WAVES=4 MODES=1 FFR_JIT_NO_DISK_CACHE=1 FFR_ACC_MAX_MB=2048
( FFR_ACC_GRAN=2 timeout 600 python -m pytest tests/test_gpu_affine.py -m gpu -x -q ) > gpurun_out/acc2_pytest.log 2>&1; tail -3 gpurun_out/acc2_pytest.log
for g in 0 2 4 7; do echo "== FFR_ACC_GRAN $g"; FFR_ACC_GRAN=$g python tools/probe.py sierpinski barnsley barnsley4k sierp4k sierp3d256 sierp3d 2>&1 | cut -c1-104; done > gpurun_out/acc2_probe.log 2>&1
cat gpurun_out/acc2_probe.log
