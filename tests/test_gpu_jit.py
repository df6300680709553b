"""K1c, the run-time compiled flame-specialised kernel (csrc/ffr_jit_kernel.cuh), against the
ahead-of-time interpreter kernels on the same seeded chains.

Both are built from the same device functions (calc2d_body / calc_nd_body / affine_apply /
ISAAC) with -fmad=false, and a chain's stream and arithmetic do not depend on which thread
advances it, so the bar is BIT-EXACT histogram counts and statistics for EVERY flame --
including the transcendental-heavy ones (class iii), where the oracle comparison can only be
statistical. Colour sums are floating point atomics (order not defined): 1e-12 relative.
The interpreter kernels themselves are checked against the oracle in test_gpu_parity.py.
"""
import numpy as np
import pytest

import flames

pytestmark = pytest.mark.gpu


def render(ffr, text, chains, chain_len, seed, last_len=0, elem_size=8, bv_limit=1 << 40, **opts):
    fl = ffr.Flame(text, elem_size=elem_size)
    r = ffr.BufferRenderer(fl, **opts)
    ok = r.render_chains(0, chains, chain_len, last_len=last_len, base_seed=seed, bv_limit=bv_limit)
    buf = r.read_buffer()
    st = r.stats
    info = r.jit_info
    r.close()
    return fl, buf, st, ok, info


def compare(ffr, text, chains=700, chain_len=300, seed=5, last_len=77, elem_size=8, **kw):
    fl, b0, s0, ok0, i0 = render(ffr, text, chains, chain_len, seed, last_len, elem_size,
                                 jit=ffr.JIT_OFF, **kw)
    fl, b1, s1, ok1, i1 = render(ffr, text, chains, chain_len, seed, last_len, elem_size,
                                 jit=ffr.JIT_ON, **kw)
    assert not i0["active"] and i1["active"], (i0, i1)
    assert ok0 == ok1
    for k in ("s_iter", "s_plot", "xf_dist", "n_bad", "pt_min", "pt_max"):
        assert s0[k] == s1[k], (k, s0[k], s1[k])
    _, _, cells, cs = fl.layout()
    c0, col0 = ffr.split_counts_colors(b0, cells, cs - 1)
    c1, col1 = ffr.split_counts_colors(b1, cells, cs - 1)
    assert np.array_equal(c0, c1)
    if cs > 1:
        np.testing.assert_allclose(col0, col1, rtol=1e-12 if elem_size == 8 else 2e-4, atol=1e-9)
    return s1, i1


EXAMPLES_2D = ["barnsley_fern", "csci6360_project", "flam3_test_1", "rectangle_maze",
               "sierpinski_triangle", "sierpinski_with_variations", "tkoz_test1", "tkoz_test2",
               "tkoz_test3", "tkoz_test4", "tkoz_test5"]


@pytest.mark.parametrize("name", EXAMPLES_2D)
def test_examples_bit_exact_vs_interpreter(ffr, examples, name):
    st, info = compare(ffr, examples.example_json(name, size=[160, 120]))
    assert st["s_iter"] == 699 * 300 + 77
    assert info["slots_per_block"] >= info["threads_per_block"] and info["registers"] > 0


def test_3d_example_and_float_build(ffr, examples):
    compare(ffr, examples.example_json("sierpinski_triangle_3d", size=[40, 40, 40]))
    compare(ffr, examples.example_json("csci6360_project", size=[160, 120]), elem_size=4)
    compare(ffr, examples.example_json("tkoz_test3", size=[160, 120]), elem_size=4)


def _groups(names, n):
    return [names[i:i + n] for i in range(0, len(names), n)]


@pytest.mark.parametrize("group", _groups(flames.ALL_VARIATIONS, 7), ids=lambda g: g[0])
def test_every_variation_2d(ffr, group):
    """All 98 variations, 7 per compiled flame, with colours and a final xform; covers the
    random-number drawing variations (draw order) and the bad value / re-init path."""
    compare(ffr, flames.multi_variation_flame(group, dims=2, final_name=group[-1]))


@pytest.mark.parametrize("group", _groups(flames.ALL_VARIATIONS[::3], 7), ids=lambda g: g[0])
def test_variations_lifted_to_3d(ffr, group):
    compare(ffr, flames.multi_variation_flame(group, dims=3, final_name=group[0]), chains=300)


def test_one_d(ffr):
    compare(ffr, flames.multi_variation_flame(list(flames.PARAMS_ND)[:7], dims=1))
    compare(ffr, flames.one_d_flame())


def test_bad_values_reinit_and_abort(ffr):
    st, _ = compare(ffr, flames.divergent_flame(), chains=600, chain_len=512)
    assert st["n_bad"] > 0
    # bad value limit: both kernels stop and report failure
    for jit in (ffr.JIT_OFF, ffr.JIT_ON):
        _, _, st, ok, _ = render(ffr, flames.divergent_flame(), 600, 512, 3, bv_limit=5, jit=jit)
        assert not ok and st["n_bad"] > 5


def test_short_and_ragged_launches(ffr, examples):
    text = examples.example_json("csci6360_project", size=[96, 54])
    compare(ffr, text, chains=1, chain_len=256, last_len=0)       # one chain
    compare(ffr, text, chains=513, chain_len=257, last_len=1)     # one slot past a block, 1-sample tail
    compare(ffr, text, chains=40000, chain_len=256, last_len=0)   # several groups per block


def test_unsupported_flame_fails_loudly(ffr):
    with pytest.raises(ffr.FfrError):
        ffr.BufferRenderer(ffr.Flame(flames.many_xforms_flame(20)), jit=ffr.JIT_ON)


def test_auto_mode_is_lazy(ffr, examples, monkeypatch):
    text = examples.example_json("csci6360_project", size=[100, 50])   # a size no other test compiles
    monkeypatch.setenv("FFR_JIT_NO_DISK_CACHE", "1")
    r = ffr.BufferRenderer(ffr.Flame(text))
    r.render(100000, 1000)
    assert not r.jit_info["active"]     # small render: interpreter kernel, nothing compiled
    b0 = r.read_buffer()
    r.jit_enable()
    assert r.jit_info["active"] and "jx_0" in r.jit_source
    r.close()
    # the cubin is in the process cache now: auto mode takes it even for a small render
    monkeypatch.setenv("FFR_JIT_USE_CACHED", "1")
    r = ffr.BufferRenderer(ffr.Flame(text))
    r.render(100000, 1000)
    assert r.jit_info["active"] and r.jit_info["from_cache"]
    assert np.array_equal(r.read_buffer(), b0)
    r.close()
    # ... unless told not to
    monkeypatch.setenv("FFR_JIT_USE_CACHED", "0")
    r = ffr.BufferRenderer(ffr.Flame(text))
    r.render(100000, 1000)
    assert not r.jit_info["active"]
    r.close()


def test_few_very_long_chains_do_not_trip_the_watchdog(ffr, po, examples):
    """with far fewer live chains than threads, almost every warp of K1d finds nothing
    to pop for the whole launch (seconds). That is waiting, not a stall: the watchdog only counts
    polls during which NO warp of the block popped anything. 6 chains x 1.5e6 samples of an
    IEEE-only flame (spherical: class ii), so the result is also bit-exact against the oracle."""
    fl = ffr.Flame(examples.example_json("flam3_test_1", size=[96, 96]))
    r = ffr.BufferRenderer(fl, jit=ffr.JIT_ON)
    assert "K1d queue-scheduled kernel" in r.jit_info["message"]
    assert r.render_chains(0, 6, 1_500_000, base_seed=77, bv_limit=1 << 40)
    got, st = r.read_buffer(), r.stats
    r.close()
    want, ost, _ = po.oracle_render(fl, 6, 1_500_000, base_seed=77, bv_limit=1 << 40, nthreads=6)
    assert st["s_iter"] == 9_000_000
    for k in ("s_iter", "s_plot", "xf_dist", "n_bad", "pt_min", "pt_max"):
        assert st[k] == ost[k], k
    assert np.array_equal(got, want)


def test_queue_soak_many_launches_bit_exact(ffr, examples):
    """Soak of K1d's slot queues (release/acquire ring entries, lap parity, predicated one-lane
    atomics): >= 100 launches of ragged chain ranges, 1e11 samples in all (FFR_SOAK_SAMPLES
    overrides; profiles/ holds a 1e12 run), on an IEEE-only flame (spherical: class ii), so every
    count and statistic must equal the ahead-of-time kernel's bit for bit. A lost or duplicated
    slot, or a slot whose state was read before it was visible, changes counts."""
    import os
    total = int(float(os.environ.get("FFR_SOAK_SAMPLES", "1e11")))
    launches, L = 100, 4096
    per = total // launches // L
    fl = ffr.Flame(examples.example_json("flam3_test_1", size=[512, 512]))
    res = []
    for jit in (ffr.JIT_ON, ffr.JIT_OFF):
        r = ffr.BufferRenderer(fl, jit=jit)
        if jit == ffr.JIT_ON:
            assert "K1d queue-scheduled kernel" in r.jit_info["message"]
        first = 0
        for k in range(launches):
            count = per + (k * 37) % 101          # ragged: a different remainder every launch
            r.render_chains_async(first, count, L, last_len=(k * 13) % L, base_seed=3, bv_limit=1 << 60)
            first += count
        r.sync()
        st = r.fetch_stats()
        res.append((r.read_buffer(), st))
        r.close()
    (b1, s1), (b0, s0) = res
    assert s1["s_iter"] >= total * 0.99
    for k in ("s_iter", "s_plot", "xf_dist", "n_bad", "pt_min", "pt_max"):
        assert s1[k] == s0[k], k
    assert np.array_equal(b1, b0)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [17, 99])
def test_polar_unit_shared_reciprocal_is_the_compilers_division(ffr, seed):
    """K1d's polar unit computes y/r and x/r from ONE refined reciprocal (div_pair in
    csrc/ffr_device.cuh: the compiler's own division sequence with the reciprocal shared, each
    quotient falling back to the compiler's division when its acceptance test fails); the
    ahead-of-time kernels divide plainly. On a flame of IEEE-only variations that read y/r and x/r
    (hyperbolic: the only class-ii variation that does) every count
    and statistic must therefore agree bit for bit -- one quotient rounded differently would send
    its chain elsewhere. 2e10 samples by default (~4e10 divisions; FFR_DIV_SOAK_SAMPLES overrides)."""
    import os
    total = int(float(os.environ.get("FFR_DIV_SOAK_SAMPLES", "2e10")))
    L = 4096
    fl = ffr.Flame(flames.variation_flame("hyperbolic", size=[512, 512], color=False, final=True))
    res = []
    for jit in (ffr.JIT_ON, ffr.JIT_OFF):
        r = ffr.BufferRenderer(fl, jit=jit)
        if jit == ffr.JIT_ON:
            assert "K1d queue-scheduled kernel" in r.jit_info["message"]
            assert "polar_fill_need" in r.jit_source
        r.render_chains(0, total // L, L, base_seed=seed, bv_limit=1 << 60)
        res.append((r.read_buffer(), r.stats))
        r.close()
    (b1, s1), (b0, s0) = res
    assert s1["s_iter"] == (total // L) * L
    for k in ("s_iter", "s_plot", "xf_dist", "n_bad", "pt_min", "pt_max"):
        assert s1[k] == s0[k], k
    assert np.array_equal(b1, b0)
