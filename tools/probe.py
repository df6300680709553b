"""Quick throughput probe (development aid, not the contract bench): wall clock around
render_chains for the BASELINE configs. Usage: python tools/probe.py [name ...]"""
import importlib, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ffr = importlib.import_module("flame-fractal-renderer_b200")
if os.environ.get("FFR_LIB"):  # development only: probe an experimental build of the library
    ffr.LIB_PATH = os.environ["FFR_LIB"]
ex = importlib.import_module("flame-fractal-renderer_b200.examples")

CONFIGS = {
    "sierpinski": ("sierpinski_triangle", [1024, 1024]),
    "barnsley": ("barnsley_fern", [2048, 2048]),
    "tkoz3": ("tkoz_test3", [4096, 4096]),
    "sierp3d": ("sierpinski_triangle_3d", [512, 512, 512]),
    "csci": ("csci6360_project", [4096, 4096]),
    "csci8k": ("csci6360_project", [8192, 8192]),
    "barnsley4k": ("barnsley_fern", [4096, 4096]),
    "barnsley1k": ("barnsley_fern", [1024, 1024]),
    "sierp4k": ("sierpinski_triangle", [4096, 4096]),
    "sierp3d256": ("sierpinski_triangle_3d", [256, 256, 256]),
    "tkoz1": ("tkoz_test1", [4096, 4096]),
    "tkoz2": ("tkoz_test2", [4096, 4096]),
    "tkoz4": ("tkoz_test4", [4096, 4096]),
    "tkoz5": ("tkoz_test5", [4096, 4096]),
    "swv": ("sierpinski_with_variations", [4096, 4096]),
    "flam3": ("flam3_test_1", [4096, 4096]),
}

def main():
    names = sys.argv[1:] or list(CONFIGS)
    L = int(os.environ.get("L", "8192"))
    waves = int(os.environ.get("WAVES", "2"))
    modes = [int(m) for m in os.environ.get("MODES", "1").split(",")]
    bps = int(os.environ.get("BPS", "0"))
    regs = [int(m) for m in os.environ.get("REGROUP", "0").split(",")]
    jits = [int(m) for m in os.environ.get("JIT", "1").split(",")]   # 1 off, 2 on
    for nm in names:
        ename, size = CONFIGS[nm]
        fl = ffr.Flame(ex.example_json(ename, size=size), elem_size=int(os.environ.get("ELEM", "8")))
        for mode, rg, jit in [(m, g, j) for m in modes for g in regs for j in jits]:
            t0 = time.time()
            r = ffr.BufferRenderer(fl, scatter_mode=mode, blocks_per_sm=bps, regroup=rg, jit=jit)
            tc = time.time() - t0
            chains = r.resident_chains * waves
            r.render_chains(0, 148 * 2 * 256, 256)  # warm-up
            t0 = time.time()
            r.render_chains(0, chains, L, base_seed=5)
            dt = time.time() - t0
            st = r.stats
            n = chains * L
            ji = r.jit_info
            print("%-10s mode %d rg %d jit %d  %.3e samples in %.3fs = %.3e samples/s  plotted %.3f  "
                  "[create %.2fs regs %d slots %d bps %d] %s" % (
                nm, mode, rg, jit, n, dt, n / dt, st["s_plot"] / st["s_iter"], tc, ji["registers"],
                ji["slots_per_block"], ji["blocks_per_sm"], ji["message"]), flush=True)
            r.close()

if __name__ == "__main__":
    main()
