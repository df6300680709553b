"""K1e, the flame-specialised kernel for pure-affine flames (csrc/ffr_jit_affine.cuh +
generate_affine in csrc/ffr_jit_host.cuh), through the C ABI.

Bar: BIT-EXACT histogram counts and statistics against (a) the oracle directly and (b) the
interpreter kernel K1 on the same seeded chains -- this is integer work on top of IEEE
+ - * only (SURVEY Q6 class i). The generated arithmetic itself is also checked on the CPU
(tests/test_affine_simplifier.py); here the whole kernel runs: seeding, the packed xform
selections, settle iterations, in-bounds / bad value / extremes bookkeeping, the index and
the scatter, ragged and long chains, the bad value re-init and abort path, the float build.
"""
import numpy as np
import pytest

import flames
from test_gpu_jit import compare, render
from test_gpu_parity import assert_buffers, assert_stats_equal

pytestmark = pytest.mark.gpu

AFFINE_EXAMPLES = ["sierpinski_triangle", "barnsley_fern", "rectangle", "rectangle_grid",
                   "sierpinski_triangle_3d"]


def is_k1e(info):
    return info["active"] and "K1e pure-affine kernel" in info["message"]


@pytest.mark.parametrize("name", AFFINE_EXAMPLES)
def test_examples_vs_oracle(ffr, po, examples, name):
    size = [48, 40, 56] if name.endswith("3d") else [200, 150]
    text = examples.example_json(name, size=size)
    fl = ffr.Flame(text)
    r = ffr.BufferRenderer(fl, jit=ffr.JIT_ON)
    assert is_k1e(r.jit_info), r.jit_info
    ok = r.render_chains(0, 1000, 700, last_len=123, base_seed=11, bv_limit=256)
    gbuf, gst = r.read_buffer(), r.stats
    r.close()
    obuf, ost, ook = po.oracle_render(fl, 1000, 700, base_seed=11, last_len=123, bv_limit=256, nthreads=8)
    assert ok and ook
    assert_stats_equal(gst, ost)
    assert_buffers(ffr, fl, gbuf, obuf)
    assert gst["s_iter"] == 999 * 700 + 123


@pytest.mark.parametrize("name", AFFINE_EXAMPLES)
def test_examples_vs_interpreter(ffr, examples, name):
    size = [40, 40, 40] if name.endswith("3d") else [160, 120]
    _, info = compare(ffr, examples.example_json(name, size=size))
    assert is_k1e(info), info


SHAPES = {
    "general_pre_post": dict(pre="general", post="general", nx=5),
    "perm_two_vars": dict(pre="perm", post="scale", weights=(0.75, -1.0)),
    "negzero_pre": dict(pre="scale", post="identity", neg_zero=("pre", 1, 1)),
    "negzero_post": dict(pre="general", post="perm", neg_zero=("post", 0, 2)),
    "eight_xforms": dict(pre="general", post="identity", nx=8),
    "one_xform": dict(pre="scale", post="identity", nx=1),
    "one_d": dict(dims=1, pre="general", post="scale"),
    "3d_no_post": dict(dims=3, pre="scale", post="none", nx=4),
    "3d_no_pre_vary_w": dict(dims=3, pre="none", post="general", nx=4, weights=(0.8, 0.3), vary_weights=True),
    "3d_neither": dict(dims=3, pre="none", post="none", weights=(0.5,)),
    # most samples fall outside these bounds: the not-plotted path and its bad value test
    "tight_bounds": dict(pre="general", post="identity", nx=4, bounds=[[-0.1, 0.2], [0.0, 0.3]]),
}


@pytest.mark.parametrize("shape", sorted(SHAPES))
def test_coefficient_shapes(ffr, shape):
    st, info = compare(ffr, flames.affine_flame(**SHAPES[shape]))
    assert is_k1e(info), info
    if shape == "tight_bounds":
        assert 0 < st["s_plot"] < st["s_iter"] // 2


def test_float_build(ffr, examples):
    for text in (examples.example_json("barnsley_fern", size=[160, 120]),
                 examples.example_json("sierpinski_triangle_3d", size=[40, 40, 40]),
                 flames.affine_flame(pre="general", post="perm", nx=5, weights=(0.5, 0.25))):
        _, info = compare(ffr, text, elem_size=4)
        assert is_k1e(info), info


def expanding_flame():
    """Pure-affine and K1e-eligible (no colours, no final xform), with one expanding xform:
    bad values (|x| > 1e20), re-initialisation, the bad value limit."""
    import json
    return json.dumps({
        "dimensions": 2, "size": [128, 128], "bounds": [[-2, 2], [-2, 2]],
        "xforms": [
            {"weight": 1.0, "variations": [{"name": "linear", "weight": 1.0}],
             "pre_affine": {"A": [[0.5, 0.0], [0.0, 0.5]], "b": [0.25, -0.25]}},
            {"weight": 0.9, "variations": [{"name": "linear", "weight": 1.0}],
             "pre_affine": {"A": [[1.0e4, 0.0], [0.0, -1.0e4]], "b": [0.1, 0.0]}},
        ]})


def test_bad_values_reinit_and_abort(ffr, po):
    text = expanding_flame()
    st, info = compare(ffr, text, chains=600, chain_len=512)
    assert is_k1e(info) and st["n_bad"] > 0
    # and against the oracle: the re-init draws continue the chain's own stream
    fl = ffr.Flame(text)
    r = ffr.BufferRenderer(fl, jit=ffr.JIT_ON)
    ok = r.render_chains(0, 300, 400, base_seed=9, bv_limit=1 << 40)
    gbuf, gst = r.read_buffer(), r.stats
    r.close()
    obuf, ost, ook = po.oracle_render(fl, 300, 400, base_seed=9, bv_limit=1 << 40, nthreads=8)
    assert ok and ook
    assert_stats_equal(gst, ost)
    assert_buffers(ffr, fl, gbuf, obuf)
    # bad value limit: the kernel stops and the call reports failure
    _, _, st, ok, _ = render(ffr, text, 600, 512, 3, bv_limit=5, jit=ffr.JIT_ON)
    assert not ok and st["n_bad"] > 5


def test_ragged_and_long_chains(ffr, examples):
    text = examples.example_json("barnsley_fern", size=[96, 54])
    compare(ffr, text, chains=1, chain_len=256, last_len=0)        # one chain
    compare(ffr, text, chains=33, chain_len=257, last_len=1)       # one lane past a warp, 1-sample tail
    compare(ffr, text, chains=100000, chain_len=256, last_len=0)   # many groups per warp
    compare(ffr, text, chains=70, chain_len=70001, last_len=40000)  # xf_dist fields flushed mid-chain


def test_scatter_modes(ffr, examples):
    """warp-aggregated / discard scatter run through K1e's second entry point"""
    text = examples.example_json("sierpinski_triangle", size=[64, 64])
    fl, b0, s0, _, _ = render(ffr, text, 500, 300, 5, jit=ffr.JIT_ON)
    fl, b1, s1, _, i1 = render(ffr, text, 500, 300, 5, jit=ffr.JIT_ON, scatter_mode=ffr.SCATTER_WARP_AGG)
    assert is_k1e(i1)
    assert np.array_equal(b0, b1) and s0["s_plot"] == s1["s_plot"]
    fl, b2, s2, _, _ = render(ffr, text, 500, 300, 5, jit=ffr.JIT_ON, scatter_mode=ffr.SCATTER_DISCARD)
    assert not b2.any() and s2["s_plot"] == s0["s_plot"]


def test_flames_outside_k1e_keep_working(ffr):
    """a pure-affine flame with colours and a final xform: JIT_ON compiles the general
    flame-specialised kernel instead, same bits"""
    st, info = compare(ffr, flames.divergent_flame(), chains=300, chain_len=300)
    assert info["active"] and not is_k1e(info)


@pytest.fixture
def force_compact_tile(monkeypatch):
    """The compact tile (row directory) is meant for buffers beyond 256 MiB; force it onto the
    small test buffers. The settings are read when the kernel is generated."""
    def setter(tile_mb=192):
        monkeypatch.setenv("FFR_DIR_MIN_MB", "0")
        monkeypatch.setenv("FFR_K1E_SCRAMBLE", "0")
        monkeypatch.setenv("FFR_DIR_TILE_MB", str(tile_mb))
    return setter


def test_compact_tile(ffr, examples, force_compact_tile):
    force_compact_tile()
    for text in (examples.example_json("sierpinski_triangle_3d", size=[64, 64, 64]),
                 examples.example_json("barnsley_fern", size=[512, 512])):
        _, info = compare(ffr, text, chains=2000, chain_len=500)
        assert is_k1e(info) and "compact tile" in info["message"], info


def test_compact_tile_overflow_and_repeated_renders(ffr, examples, force_compact_tile):
    """1 MiB tile = 256 rows of 512 cells; the fern at 1024x512 touches more rows than that: the
    rest is scattered into the buffer directly. Rows keep their slots across render calls."""
    force_compact_tile(tile_mb=1)
    text = examples.example_json("barnsley_fern", size=[1024, 512])
    fl = ffr.Flame(text)
    bufs = []
    for jit in (ffr.JIT_OFF, ffr.JIT_ON):
        r = ffr.BufferRenderer(fl, jit=jit)
        if jit == ffr.JIT_ON:
            assert "compact tile of 256 rows" in r.jit_info["message"], r.jit_info
        for k in range(3):
            assert r.render_chains(1000 * k, 1000, 400, base_seed=3)
        bufs.append((r.read_buffer(), r.stats))
        r.close()
    assert np.array_equal(bufs[0][0], bufs[1][0])
    for k in ("s_iter", "s_plot", "xf_dist", "pt_min", "pt_max"):
        assert bufs[0][1][k] == bufs[1][1][k], k
    assert np.count_nonzero(bufs[1][0].reshape(-1, 512).any(axis=1)) > 256   # it did overflow


def test_cli_jit_affine_vs_oracle(ffr, po, examples, tmp_path):
    """ffr-buf.out --jit on a pure-affine flame runs K1e (accumulation tile + fold): the buffer
    file equals the oracle's, and -i still adds on top of it."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(ffr.LIB_PATH), "ffr-buf.out")
    flame = tmp_path / "fern.json"
    flame.write_text(examples.example_json("barnsley_fern", size=[256, 128]))
    out1, out2 = tmp_path / "a.buf", tmp_path / "b.buf"
    p = subprocess.run([exe, "-f", str(flame), "-o", str(out1), "-s", "1500000", "-b", "1000",
                        "--seed", "5", "--jit"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    a = np.fromfile(out1, dtype=np.uint64)
    fl = ffr.Flame(flame.read_text())
    want, _, _ = po.oracle_render_samples(fl, 1_500_000, 1000, base_seed=5, nthreads=8)
    assert np.array_equal(a, want)
    p = subprocess.run([exe, "-f", str(flame), "-i", str(out1), "-o", str(out2), "-s", "1500000",
                        "-b", "1000", "--seed", "5", "--jit"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert np.array_equal(np.fromfile(out2, dtype=np.uint64), 2 * a)


def test_two_devices_k1e(ffr, po, examples):
    """K1e on a two-device context: a tile per device, folded before the peer-memory reduce."""
    if ffr.lib().ffr_cuda_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    fl = ffr.Flame(examples.example_json("sierpinski_triangle", size=[256, 256]))
    r = ffr.BufferRenderer(fl, devices=[0, 1], jit=ffr.JIT_ON)
    assert is_k1e(r.jit_info)
    calls = []
    assert r.render(3_000_000, 1000, base_seed=8, progress=lambda d, t: calls.append((d, t)))
    got, st = r.read_buffer(), r.stats
    r.close()
    want, ost, _ = po.oracle_render_samples(fl, 3_000_000, 1000, base_seed=8, nthreads=8)
    assert np.array_equal(got, want)
    assert_stats_equal(st, ost)
    assert calls and calls[-1][0] == calls[-1][1]
