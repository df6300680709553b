"""f2: the float/uint32_t build (num_t = float, hist_t = uint32_t, ISAAC-32, settle 24, eps
1e-10, bad value 1e10; reference types/types.hpp:24-41) on the device, checked DIRECTLY against
the reference compiled in that configuration (oracle/_ref/libffr_ref_f32.so, `make -C oracle
ref32`; there is no separate restatement for this build). Skipped when that library is absent.

The float reference is a mixed-precision program: unqualified sin/cos/atan2/exp/log/pow resolve
to the double libm functions on promoted arguments, sincosg to sincosf, sqrt to the double sqrt.
The device code keeps exactly those promotions, so the bars are the same as for the double build:
bit-exact counts/statistics for pure-affine and IEEE-only flames; one application within 2e-6
relative (a few float ULP) for every variation; statistical parity for variation-heavy renders.
"""
import numpy as np
import pytest

import flames
import pyoracle

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not pyoracle.have_ref(4), reason="oracle/_ref/libffr_ref_f32.so not built")]

EXACT_EXAMPLES = ["sierpinski_triangle", "barnsley_fern", "rectangle", "rectangle_grid",
                  "sierpinski_triangle_3d", "flam3_test_1", "flam3_test_2"]
REGROUP = {"direct": 1, "regroup": 2}


def both(ffr, po, text, chains, L, seed=1, last_len=0, bv_limit=1 << 40, **opts):
    fl = ffr.Flame(text, elem_size=4)
    r = ffr.BufferRenderer(fl, **opts)
    assert r.elem_size == 4 and r.bytes == fl.layout()[2] * fl.layout()[3] * 4
    ok = r.render_chains(0, chains, L, last_len=last_len, base_seed=seed, bv_limit=bv_limit)
    g, gst = r.read_buffer(), r.stats
    r.close()
    o, ost, ook = po.ref_render(text, chains, L, base_seed=seed, last_len=last_len,
                                bv_limit=1 << 20, elem_size=4)
    return fl, g, gst, ok, o, ost, ook


def check_exact(ffr, fl, g, gst, o, ost):
    _, _, cells, cs = fl.layout()
    gc, gcol = ffr.split_counts_colors(g, cells, cs - 1)
    oc, ocol = ffr.split_counts_colors(o, cells, cs - 1)
    assert gc.dtype == np.uint32
    assert np.array_equal(gc, oc)
    for k in ("s_iter", "s_plot", "xf_dist", "n_bad", "pt_min", "pt_max"):
        assert gst[k] == ost[k], k
    if cs > 1:  # float atomic adds in a different order
        np.testing.assert_allclose(gcol, ocol, rtol=2e-4, atol=1e-3)


def test_isaac32_stream_bit_exact(ffr, po, examples):
    fl = ffr.Flame(examples.example_json("sierpinski_triangle"), elem_size=4)
    r = ffr.BufferRenderer(fl)
    for seed in (1, 2, 12345, 2**64 - 1, 0xDEADBEEF12345678):
        assert np.array_equal(r.isaac_words(seed, 100), po.ref_isaac_words(seed, 100, elem_size=4)), seed
    r.close()


@pytest.mark.parametrize("kernel", sorted(REGROUP))
@pytest.mark.parametrize("name", EXACT_EXAMPLES)
def test_exact_examples_f32(ffr, po, examples, name, kernel):
    size = [48, 48, 48] if name.endswith("3d") else None
    text = examples.example_json(name, size=size)
    fl, g, gst, ok, o, ost, ook = both(ffr, po, text, 400, 700, seed=11, last_len=123,
                                       regroup=REGROUP[kernel])
    assert ok and ook
    check_exact(ffr, fl, g, gst, o, ost)


@pytest.mark.parametrize("name", flames.IEEE_EXACT)
@pytest.mark.parametrize("dims", [2, 3])
def test_ieee_only_variations_bit_exact_f32(ffr, po, name, dims):
    if dims == 3 and name not in flames.PARAMS_ND and name not in ("horseshoe", "curl", "boarders"):
        pytest.skip("3-d lifting covered by a subset")
    text = flames.variation_flame(name, dims=dims, final=(dims == 3))
    fl, g, gst, ok, o, ost, ook = both(ffr, po, text, 200, 512, seed=3)
    check_exact(ffr, fl, g, gst, o, ost)


def test_edge_flames_f32(ffr, po):
    for text in (flames.one_d_flame(), flames.many_xforms_flame()):
        fl, g, gst, ok, o, ost, ook = both(ffr, po, text, 200, 300, seed=21)
        check_exact(ffr, fl, g, gst, o, ost)


def test_bad_values_f32(ffr, po):
    """Every iteration of this flame ends in a bad value (re-init path). The stale
    pf it then tests is NaN: the reference's undefined NaN -> size_t cast happens to count it
    in cell 0 on x86-64, the device fences NaN as out of bounds. Everything the
    fence does not touch must still match the float reference exactly."""
    fl, g, gst, ok, o, ost, ook = both(ffr, po, flames.divergent_flame(), 200, 300, seed=21)
    for k in ("s_iter", "xf_dist", "n_bad", "pt_min", "pt_max"):
        assert gst[k] == ost[k], k
    assert gst["n_bad"] > 0.9 * 200 * 300
    _, _, cells, cs = fl.layout()
    gc, _ = ffr.split_counts_colors(g, cells, cs - 1)
    oc, _ = ffr.split_counts_colors(o, cells, cs - 1)
    assert np.array_equal(gc[1:], oc[1:])               # only cell 0 holds the reference's NaN hits
    assert int(oc[0]) - int(gc[0]) == ost["s_plot"] - gst["s_plot"]


@pytest.mark.parametrize("name", flames.ALL_VARIATIONS)
def test_every_variation_single_step_f32(ffr, po, name):
    for dims in ((1, 2, 3) if name in flames.PARAMS_ND else (2, 3)):
        text = flames.variation_flame(name, dims=dims, final=True)
        fl = ffr.Flame(text, elem_size=4)
        r = ffr.BufferRenderer(fl)
        rng = np.random.default_rng(99 + dims)
        n = 2048
        pts = rng.uniform(-2.5, 2.5, size=(n, dims)).astype(np.float32).astype(np.float64)
        seeds = rng.integers(0, 2**63, size=n, dtype=np.uint64)
        for xi in list(range(fl.desc.num_xforms)) + [-1]:
            got = r.iterate_points(xi, seeds, pts)
            want = po.ref_iterate_points(text, xi, seeds, pts, elem_size=4)
            if name in flames.IEEE_EXACT or name == "linear":
                assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), (dims, xi)
            else:
                fin = np.isfinite(want).all(axis=1) & (np.abs(want) < 1e6).all(axis=1)
                assert np.array_equal(np.isnan(got), np.isnan(want)), (dims, xi)
                scale = np.maximum(1.0, np.abs(want[fin]).max(axis=1, keepdims=True))
                err = np.abs(got[fin] - want[fin]) / scale
                frac_bad = float((err > 2e-6).any(axis=1).mean())
                assert frac_bad < 5e-3, (dims, xi, frac_bad, float(err.max()))
        r.close()


def _coarse(c, size, f=4):
    w, h = size
    return c.reshape(h, w)[: h - h % f, : w - w % f].astype(np.float64).reshape(h // f, f, w // f, f).sum(axis=(1, 3)).ravel()


@pytest.mark.parametrize("name,size", [("csci6360_project", [192, 108]), ("tkoz_test3", [160, 90])])
def test_statistical_parity_f32(ffr, po, examples, name, size):
    """Same measured-noise-floor criterion as the double build (test_gpu_parity.py)."""
    text = examples.example_json(name, size=size)
    chains, L = 4096, 512
    fl = ffr.Flame(text, elem_size=4)
    _, _, cells, cs = fl.layout()
    def l1(a, b):
        return float(np.abs(a / a.sum() - b / b.sum()).sum())
    g = []
    for seed in (1, 90001):
        r = ffr.BufferRenderer(fl)
        assert r.render_chains(0, chains, L, base_seed=seed)
        gc, _ = ffr.split_counts_colors(r.read_buffer(), cells, cs - 1)
        st = r.stats
        r.close()
        assert int(gc.sum()) == st["s_plot"] and st["s_iter"] == chains * L
        g.append((_coarse(gc, size), st))
    o = []
    for seed in (100_001, 200_777, 304_242):   # further apart than the chain count
        ob, ost, _ = po.ref_render(text, chains, L, base_seed=seed, elem_size=4)
        oc, _ = ffr.split_counts_colors(ob, cells, cs - 1)
        o.append((_coarse(oc, size), ost))
    pairs = [(0, 1), (0, 2), (1, 2)]
    floor = max(l1(o[i][0], o[j][0]) for i, j in pairs)
    worst = max(l1(a[0], b[0]) for a in g for b in o)
    assert worst <= 1.5 * floor, (worst, floor)
    oplot = [x[1]["s_plot"] for x in o]
    n = chains * L
    p = np.mean(oplot) / n
    tol = 1.5 * (max(oplot) - min(oplot)) + 5 * (n * p * (1 - p)) ** 0.5 + 1
    for a in g:
        assert min(oplot) - tol <= a[1]["s_plot"] <= max(oplot) + tol


def test_f32_buffer_interop_and_tonemap(ffr, examples):
    """4-byte buffers: add_buffer doubles counts and colour sums, the tone map runs in float."""
    fl = ffr.Flame(examples.example_json("tkoz_test3", size=[96, 54]), elem_size=4)
    _, _, cells, cs = fl.layout()
    r = ffr.BufferRenderer(fl)
    r.render_chains(0, 300, 512, base_seed=3)
    first = r.read_buffer().copy()
    assert first.dtype == np.uint32 and first.nbytes == cells * cs * 4
    r.add_buffer(first)
    twice = r.read_buffer()
    c1, col1 = ffr.split_counts_colors(first, cells, cs - 1)
    c2, col2 = ffr.split_counts_colors(twice, cells, cs - 1)
    assert np.array_equal(c2, 2 * c1)
    np.testing.assert_array_equal(col2, col1 + col1)
    s, m = r.histogram_sum_max()
    assert s == int(c2.sum()) and m == int(c2.max())
    img, info = r.tonemap(ffr.TONE_GRAY, bits=8, gamma=2.0)
    # the same pixel math in float32 on the host (ffr_img.cpp:236-243 with num_t = float)
    n = c2.reshape(54, 96).astype(np.float32)
    l = np.log(np.float32(1) + n) / np.log(np.float32(1) + np.float32(c2.max()))
    want = np.minimum(np.power(l, np.float32(0.5)) * (np.float32(256) * (np.float32(1) - np.float32(2) ** -23)), 255)
    d = np.abs(img.astype(np.int64) - want.astype(np.uint8).astype(np.int64))
    assert d.max() <= 1 and (d != 0).mean() < 0.02
    r.close()


def test_cli_float_build(ffr, po, examples, tmp_path):
    """ffr-buf.out --float writes 4-byte elements identical to the float reference; ffr-img.out
    --float reads them."""
    import os
    import subprocess
    here = os.path.dirname(ffr.LIB_PATH)
    flame = tmp_path / "fern.json"
    text = examples.example_json("barnsley_fern", size=[160, 100])
    flame.write_text(text)
    out = tmp_path / "f.buf"
    p = subprocess.run([os.path.join(here, "ffr-buf.out"), "-f", str(flame), "-o", str(out), "-s", "300000",
                        "-b", "1000", "--seed", "5", "--float"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert "4 byte numbers" in p.stderr
    a = np.fromfile(out, dtype=np.uint32)
    want, _, _ = po.ref_render(text, 300, 1000, base_seed=5, elem_size=4)
    assert np.array_equal(a, want)
    png = tmp_path / "f.png"
    p = subprocess.run([os.path.join(here, "ffr-img.out"), "-f", str(flame), "-i", str(out), "-o", str(png),
                        "-g", "--float"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert open(png, "rb").read(8) == b"\x89PNG\r\n\x1a\n"
