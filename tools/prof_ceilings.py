"""The scatter-ceiling microbenchmarks alone, for ncu (development aid): uniform random REDs,
streamed attractor replay and windowed attractor replay on one workload.
Usage: prof_ceilings.py workload"""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ffr = importlib.import_module("flame-fractal-renderer_b200")
ex = importlib.import_module("flame-fractal-renderer_b200.examples")
from probe import CONFIGS
ename, size = CONFIGS[sys.argv[1]]
r = ffr.BufferRenderer(ffr.Flame(ex.example_json(ename, size=size)), jit=2)
for pattern, n in ((0, 1 << 28), (1, 1 << 27), (2, 1 << 28)):
    r.atomic_roofline(1 << 24, pattern=pattern)
    ms, cells = r.atomic_roofline(n, pattern=pattern)
    print("pattern %d: %.3e cells/s" % (pattern, cells / (ms * 1e-3)), flush=True)
r.close()
