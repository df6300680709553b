mkdir -p gpurun_out
export FFR_JIT_NO_DISK_CACHE=1
python - > gpurun_out/dir3_probe.log 2>&1 <<'PY'
import sys, os, time, importlib
sys.path.insert(0, os.getcwd())
ffr = importlib.import_module("flame-fractal-renderer_b200")
ex = importlib.import_module("flame-fractal-renderer_b200.examples")
CFG = [("sierpinski_triangle_3d", [512,512,512]), ("barnsley_fern", [8192,8192]), ("sierpinski_triangle", [8192,8192])]
for name, size in CFG:
    for scr in (0,1,2,3):
        os.environ["FFR_EXPERIMENT_ROWSCR"] = str(scr)
        fl = ffr.Flame(ex.example_json(name, size=size))
        r = ffr.BufferRenderer(fl, jit=2)
        chains = 148*768*4
        r.render_chains(0, 148*3*256, 256)
        for rep in range(2):
            t0 = time.time(); r.render_chains(0, chains, 8192, base_seed=5 + rep); dt = time.time() - t0
        print("%-24s %-16s rowscr %d  %.3e samples/s" % (name, size, scr, chains*8192/dt), flush=True)
        r.close()
PY
cat gpurun_out/dir3_probe.log
