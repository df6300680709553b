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
# the run-time compiled kernels: K1d (queue scheduled), K1c (lock step, FFR_JIT_ASYNC=0), K1e
# (pure-affine: scatter into the buffer, scrambled accumulation tile, compact tile behind the row
# directory forced onto a small buffer with a tile that overflows); env is read at context creation
J = {"jit": 2}
jit_cases = [
    (ex.example_json("csci6360_project", size=[64, 36]), J, {}),                               # K1d
    (ex.example_json("tkoz_test3", size=[64, 36]), J, {}),                                      # K1d colour
    (flames.variation_flame("julian", dims=3, final=True), J, {}),                              # K1d rng in smem
    (flames.divergent_flame(), J, {}),                                                          # K1d bad values
    (ex.example_json("csci6360_project", size=[64, 36]), J, {"FFR_JIT_ASYNC": "0"}),           # K1c
    (ex.example_json("tkoz_test3", size=[64, 36]), J, {"FFR_JIT_ASYNC": "0"}),                 # K1c colour
    (ex.example_json("sierpinski_triangle", size=[64, 64]), J, {}),                             # K1e + tile
    (ex.example_json("barnsley_fern", size=[100, 60]), J, {}),                                  # K1e direct
    (ex.example_json("sierpinski_triangle", size=[2048, 2048]), J, {"FFR_ACC_MAX_MB": "16"}),  # K1e direct, big
    (ex.example_json("sierpinski_triangle_3d", size=[64, 64, 64]), J,
     {"FFR_K1E_SCRAMBLE": "0", "FFR_DIR_MIN_MB": "0", "FFR_DIR_TILE_MB": "1"}),                 # K1e + directory, overflowing
    (ex.example_json("sierpinski_triangle_3d", size=[64, 64, 64]), J,
     {"FFR_K1E_SCRAMBLE": "0", "FFR_DIR_MIN_MB": "0"}),                                         # K1e + directory
]
if os.environ.get("SANITIZE_JIT", "1") != "0":
    cases = [(t, k, {}) for t, k in cases] + jit_cases
else:
    cases = [(t, k, {}) for t, k in cases]
for text, kw, env in cases:
    saved = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
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
    print(kw, env, r.jit_info["message"][:60], r.stats["s_iter"], r.stats["n_bad"], flush=True)
    r.close()
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
