mkdir -p gpurun_out
export FFR_JIT_NO_DISK_CACHE=1
python - > gpurun_out/dir2_probe.log 2>&1 <<'PY'
import sys, os, time, importlib
sys.path.insert(0, os.getcwd())
ffr = importlib.import_module("flame-fractal-renderer_b200")
ex = importlib.import_module("flame-fractal-renderer_b200.examples")
CFG = [("sierpinski_triangle_3d", [512,512,512]), ("barnsley_fern", [8192,8192])]
for name, size in CFG:
    for tpb, minb in ((256,3),(256,2),(512,1),(128,4),(256,1)):
        os.environ["FFR_JIT_TPB"] = str(tpb); os.environ["FFR_JIT_MINB"] = str(minb)
        fl = ffr.Flame(ex.example_json(name, size=size))
        r = ffr.BufferRenderer(fl, jit=2, blocks_per_sm=minb)
        chains = 148*768*4
        r.render_chains(0, 148*3*256, 256)
        for rep in range(2):
            t0 = time.time(); r.render_chains(0, chains, 8192, base_seed=5 + rep); dt = time.time() - t0
        print("%-24s %-16s tpb %d x %d  %.3e samples/s  bps %d" % (name, size, tpb, minb, chains*8192/dt, r.jit_info["blocks_per_sm"]), flush=True)
        r.close()
PY
cat gpurun_out/dir2_probe.log
