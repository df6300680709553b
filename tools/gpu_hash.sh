mkdir -p gpurun_out
export JIT=2 WAVES=4 MODES=1 FFR_JIT_NO_DISK_CACHE=1
for h in 0 1 2 3; do echo "== hash $h"; FFR_EXPERIMENT_HASH=$h python tools/probe.py sierpinski barnsley sierp3d 2>&1 | cut -c1-100; done > gpurun_out/hash_probe.log 2>&1
cat gpurun_out/hash_probe.log
