mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_affine.py -m gpu -x -q -rs ) > gpurun_out/dir_pytest.log 2>&1; tail -5 gpurun_out/dir_pytest.log
export JIT=2 WAVES=4 MODES=1 FFR_JIT_NO_DISK_CACHE=1
python - > gpurun_out/dir_probe.log 2>&1 <<'PY'
import sys, os, time, importlib
sys.path.insert(0, os.getcwd())
ffr = importlib.import_module("flame-fractal-renderer_b200")
ex = importlib.import_module("flame-fractal-renderer_b200.examples")
CFG = [("sierpinski_triangle_3d", [512,512,512]), ("barnsley_fern", [8192,8192]), ("sierpinski_triangle", [8192,8192]), ("sierpinski_triangle_3d", [1024,512,512])]
for name, size in CFG:
    for d in ("1", "0"):
        os.environ["FFR_K1E_DIR"] = d
        fl = ffr.Flame(ex.example_json(name, size=size))
        r = ffr.BufferRenderer(fl, jit=2)
        chains = r.resident_chains * 4
        r.render_chains(0, 148*3*256, 256)
        for rep in range(2):
            t0 = time.time(); r.render_chains(0, chains, 8192, base_seed=5 + rep); dt = time.time() - t0
        print("%-24s %-16s dir=%s %.3e samples/s  %s" % (name, size, d, chains*8192/dt, r.jit_info["message"][:70]), flush=True)
        r.close()
PY
cat gpurun_out/dir_probe.log
