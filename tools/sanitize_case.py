"""Small renders through every kernel family for compute-sanitizer (development aid)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
ffr = importlib.import_module("flame-fractal-renderer_b200")
ex = importlib.import_module("flame-fractal-renderer_b200.examples")
import flames
cases = [
    (ex.example_json("sierpinski_triangle", size=[64, 64]), {}),               # K1 affine
    (ex.example_json("csci6360_project", size=[64, 36]), {}),                   # K1b
    (ex.example_json("csci6360_project", size=[64, 36]), {"regroup": 1}),       # K1 general
    (ex.example_json("tkoz_test3", size=[64, 36]), {"regroup": 2}),             # K1b colour
    (flames.variation_flame("julian", dims=3, final=True), {"regroup": 2}),     # K1b rng + 3d
    (flames.divergent_flame(), {"regroup": 2}),                                 # K1b bad values
    (flames.divergent_flame(), {"regroup": 1}),
    (flames.one_d_flame(), {}),
]
for text, kw in cases:
    fl = ffr.Flame(text)
    r = ffr.BufferRenderer(fl, **kw)
    r.render_chains(0, 700, 300, last_len=77, base_seed=3, bv_limit=1 << 40)
    r.histogram_sum_max()
    if fl.dims == 2:
        try:
            r.tonemap(ffr.TONE_GRAY)
        except ffr.FfrError:
            pass
    b = r.read_buffer()
    r.add_buffer(b)
    print(kw, r.stats["s_iter"], r.stats["n_bad"], flush=True)
    r.close()
