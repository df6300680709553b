"""One render launch for ncu (development aid). Usage: prof_one.py workload regroup [waves] [L] [jit]"""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ffr = importlib.import_module("flame-fractal-renderer_b200")
ex = importlib.import_module("flame-fractal-renderer_b200.examples")
from probe import CONFIGS
nm, rg = sys.argv[1], int(sys.argv[2])
waves = int(sys.argv[3]) if len(sys.argv) > 3 else 1
L = int(sys.argv[4]) if len(sys.argv) > 4 else 2048
ename, size = CONFIGS[nm]
fl = ffr.Flame(ex.example_json(ename, size=size))
jit = int(sys.argv[5]) if len(sys.argv) > 5 else 1
r = ffr.BufferRenderer(fl, regroup=rg, jit=jit)
r.render_chains(0, 148 * 2 * 256, 256)
r.render_chains(0, r.resident_chains * waves, L, base_seed=5)
print(r.stats["s_iter"])
