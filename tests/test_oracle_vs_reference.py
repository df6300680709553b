"""Live pin of the oracle restatement against the unmodified reference (oracle/_ref), on other
seeds and sizes than the committed golden vectors. Skipped where _ref was not built."""
import numpy as np
import pytest

import flames
import pyoracle

pytestmark = pytest.mark.skipif(not pyoracle.have_ref(), reason="oracle/_ref not built")


def same(po, ffr, text, chains=16, L=1100, seed=99, last=17):
    fl = ffr.Flame(text)
    po.set_nan_emulation(True)
    try:
        b1, s1, ok1 = po.oracle_render(fl, chains, L, base_seed=seed, last_len=last, bv_limit=1 << 20)
    finally:
        po.set_nan_emulation(False)
    b2, s2, ok2 = po.ref_render(text, chains, L, base_seed=seed, last_len=last, bv_limit=1 << 20)
    assert np.array_equal(b1, b2)
    assert repr(s1) == repr(s2)
    assert ok1 == ok2


def test_isaac_stream(po):
    for seed in (0, 5, 2**63, 2**64 - 2):
        assert np.array_equal(po.oracle_isaac_words(seed, 200), po.ref_isaac_words(seed, 200))
    assert po.oracle().oracle_splitmix64(12345) == po.ref().ref_splitmix64(12345)


def test_all_examples(ffr, po, examples):
    for name in examples.EXAMPLES:
        size = [40, 40, 40] if name.endswith("3d") else None
        same(po, ffr, examples.example_json(name, size=size))


@pytest.mark.parametrize("name", flames.ALL_VARIATIONS)
def test_every_variation(ffr, po, name):
    for d in (2, 3):
        same(po, ffr, flames.variation_flame(name, dims=d, final=(d == 2)), chains=6, L=900)


def test_single_steps(ffr, po):
    rng = np.random.default_rng(5)
    for name in ("julian", "supershape", "boarders", "cpow", "lazysusan", "noise"):
        text = flames.variation_flame(name, dims=2, final=True)
        fl = ffr.Flame(text)
        pts = rng.uniform(-3, 3, size=(5000, 2))
        seeds = rng.integers(0, 2**63, size=5000, dtype=np.uint64)
        for xi in (0, 1, 2, -1):
            a = po.oracle_iterate_points(fl, xi, seeds, pts)
            b = po.ref_iterate_points(text, xi, seeds, pts)
            assert np.array_equal(a.view(np.uint64), b.view(np.uint64))


def test_reference_threaded_render_runs(ffr, po, examples):
    # the CPU baseline entry point: BufferRenderer::render with threads
    text = examples.example_json("sierpinski_triangle", size=[64, 64])
    secs, st, buf = po.ref_render_mt(text, 200_000, 2, 4096, want_buffer=True)
    assert st["s_iter"] == 200_000 and st["s_plot"] == 200_000
    assert int(buf.sum()) == 200_000 and secs > 0


@pytest.mark.skipif(not pyoracle.have_ref(4), reason="oracle/_ref/libffr_ref_f32.so not built")
def test_float_build_flatten_matches_float_reference(ffr, po, examples):
    """f2 host side: with elem_size 4 the flame model runs every constructor in float like the
    reference compiled with num_t = float (types.hpp:24-41): xform order, cumulative weights and
    index multipliers must be the float reference's, value for value."""
    texts = [examples.example_json(n) for n in examples.EXAMPLES]
    texts += [flames.variation_flame(v, dims=2, final=True) for v in flames.ALL_VARIATIONS]
    texts += [flames.many_xforms_flame(), flames.one_d_flame()]
    for text in texts:
        f = ffr.Flame(text, elem_size=4)
        info = po.ref_flame_info(text, elem_size=4)
        md, mi, cells, cs = f.layout()
        assert f.elem_size == 4
        assert f.xform_ids == info["ids"]
        assert f.cumulative_weights == info["cw"]
        assert md == info["mult_d"] and mi == info["mult_i"]
        assert cells == info["cells"] and cs == info["cell_size"]
        # every stored value is exactly a float
        for i in range(f.desc.num_xforms):
            x = f.desc.xforms[i]
            vals = list(x.pre_A) + list(x.pre_b) + [x.weight, x.color_speed]
            vals += [x.vars[k].params[q] for k in range(x.num_vars) for q in range(8)]
            assert all(float(np.float32(v)) == v for v in vals)
    # ISAAC-32 known answers of the float build
    assert list(po.ref_isaac_words(1, 4, elem_size=4)) == [676671429, 3101584658, 2918577689, 525991190]


@pytest.mark.skipif(not pyoracle.have_refimg(), reason="oracle/_ref/libffr_refimg.so not built")
@pytest.mark.parametrize("name,size", [("tkoz_test3", [200, 120]), ("tkoz_test1", [128, 96]),
                                       ("sierpinski_triangle", [64, 64])])
def test_tonemap_restatement_equals_reference_code(ffr, po, examples, name, size):
    """f1: oracle_tonemap vs the reference's own render_image() + image_renderer.hpp compiled by
    `make -C oracle refimg`, on other flames, seeds and gammas than the golden fixture."""
    text = examples.example_json(name, size=size)
    fl = ffr.Flame(text)
    raw, _, _ = po.oracle_render(fl, 500, 800, base_seed=41, nthreads=8)
    cd = fl.color_dims
    for mode in (1, 2, 3):
        if mode == 3 and cd != 3:
            with pytest.raises(RuntimeError, match="buffer must use 3 color dimensions"):
                po.ref_tonemap(text, raw, 3)
            continue
        for bits in (8, 16):
            for gamma in (1.0, 1.7, 2.2, 0.3, 1e-3):
                a, ia = po.ref_tonemap(text, raw, mode, bits, gamma)
                b, ib = po.oracle_tonemap(raw, size[0], size[1], cd, mode, bits=bits, gamma=gamma)
                assert np.array_equal(a, b), (mode, bits, gamma)
                assert ia["hist_min"] == ib["hist_min"] and ia["hist_max"] == ib["hist_max"]
    empty = np.zeros_like(raw)
    with pytest.raises(RuntimeError, match="probably. empty"):
        po.ref_tonemap(text, empty, 2)
    with pytest.raises(RuntimeError, match="gamma too small"):
        po.ref_tonemap(text, raw, 2, 8, 0.0)


@pytest.mark.skipif(not pyoracle.have_refimg(), reason="oracle/_ref/libffr_refimg.so not built")
def test_flame_echo_equals_reference_operator(ffr, po, examples):
    texts = [examples.example_json(n) for n in examples.EXAMPLES]
    texts += [flames.variation_flame(v, dims=d, final=True) for v in flames.ALL_VARIATIONS for d in (2, 3)]
    texts += [flames.many_xforms_flame(), flames.one_d_flame(), flames.divergent_flame()]
    for text in texts:
        assert ffr.flame_json_echo(text) == po.ref_flame_echo(text)
