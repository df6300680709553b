"""Parity at the BASELINE.json configurations' full buffer sizes, for BOTH kernel families: the
ahead-of-time interpreter kernels (K1/K1b) and the run-time compiled kernels the bench measures
(K1e with its cell-scrambled / sector-scrambled / compact accumulation tiles, K1d), each
against the ORACLE on the same seeded chains -- not against each other.

pure affine (bit-exact counts + statistics vs the oracle, SURVEY Q6 class i):
  cfg1 sierpinski_triangle 1024^2 (8 MiB), 1e8 samples     K1 and K1e + cell-scrambled tile
  cfg2 barnsley_fern 2048^2 (32 MiB), 1e9 samples           K1 and K1e + cell-scrambled tile
       barnsley_fern / sierpinski 4096^2 (128 MiB), 4e8     K1e + SECTOR-scrambled tile
  cfg4 sierpinski_triangle_3d 512^3 (1 GiB), 2e8            K1 and K1e + compact tile (row directory)
       barnsley_fern 8192^2 (512 MiB), 2e8, 32 MiB tile     K1e + compact tile that OVERFLOWS (rows
                                                            beyond its capacity go direct)
variation heavy (statistical vs the oracle at config size, SURVEY 8d "parity at scale"):
  cfg3 tkoz_test3 4096^2 r=3 (512 MiB), cfg5/target csci6360_project 4096^2: K1d, 2e8 samples,
  tolerance = 1.5 x the oracle's own seed-to-seed noise floor (see test_k1d_statistical_...).
Size-independent properties (checksum of checksums, coarsening, -i additivity) stay for the
largest buffers, where every GPU kernel is also compared with a small-size render.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

_ORACLE = {}


def oracle_cached(po, ffr, examples, name, size, n, L, seed):
    key = (name, tuple(size), n, L, seed)
    if key not in _ORACLE:
        _ORACLE.clear()     # one large buffer at a time
        fl = ffr.Flame(examples.example_json(name, size=size))
        want, ost, ook = po.oracle_render_samples(fl, n, L, base_seed=seed,
                                                  nthreads=min(32, os.cpu_count() or 8))
        assert ook
        _ORACLE[key] = (want, ost)
    return _ORACLE[key]


def render_exact(ffr, po, examples, name, size, n, L, seed, jit, expect):
    """One render through the C ABI, bit for bit against the oracle: counts, s_iter, s_plot,
    xform selection counts, extremes. `expect`: substring of ffr_jit_info.message naming the
    kernel variant that must have run (None: the ahead-of-time kernels)."""
    fl = ffr.Flame(examples.example_json(name, size=size))
    r = ffr.BufferRenderer(fl, jit=jit)
    info = r.jit_info
    if expect is None:
        assert not info["active"], info
    else:
        assert info["active"] and expect in info["message"], info
    assert r.render(n, L, base_seed=seed)
    got, st = r.read_buffer(), r.stats
    s, m = r.histogram_sum_max()
    r.close()
    want, ost = oracle_cached(po, ffr, examples, name, size, n, L, seed)
    assert np.array_equal(got, want)
    for k in ("s_iter", "s_plot", "xf_dist", "pt_min", "pt_max", "n_bad"):
        assert st[k] == ost[k], k
    assert st["s_iter"] == n and s == st["s_plot"] == int(got.sum()) and m == int(got.max())


KERNELS = {"aot": None, "k1e": "K1e pure-affine kernel"}


@pytest.mark.parametrize("kernel", ["aot", "k1e"])
def test_cfg1_sierpinski_1024_bit_exact(ffr, po, examples, kernel):
    expect = None if kernel == "aot" else "cell-scrambled accumulation tile"
    render_exact(ffr, po, examples, "sierpinski_triangle", [1024, 1024], 100_000_000, 8192, 1,
                 ffr.JIT_OFF if kernel == "aot" else ffr.JIT_ON, expect)


@pytest.mark.parametrize("kernel", ["aot", "k1e"])
def test_cfg2_barnsley_2048_1e9_bit_exact(ffr, po, examples, kernel):
    expect = None if kernel == "aot" else "cell-scrambled accumulation tile"
    render_exact(ffr, po, examples, "barnsley_fern", [2048, 2048], 1_000_000_000, 8192, 1,
                 ffr.JIT_OFF if kernel == "aot" else ffr.JIT_ON, expect)


@pytest.mark.parametrize("name", ["barnsley_fern", "sierpinski_triangle"])
def test_128mib_sector_scrambled_tile_bit_exact(ffr, po, examples, name):
    """4096^2 counts = 128 MiB: the tile scrambles whole 32-byte sectors (cells stay together)."""
    render_exact(ffr, po, examples, name, [4096, 4096], 400_000_000, 8192, 5, ffr.JIT_ON,
                 "sector-scrambled accumulation tile")


@pytest.mark.parametrize("kernel", ["aot", "k1e"])
def test_cfg4_sierpinski3d_512_bit_exact(ffr, po, examples, kernel):
    """1 GiB: K1e scatters into the compact tile behind the row directory (first-touch row
    allocation by CAS, directory read through L1), folded into the buffer by K2c."""
    expect = None if kernel == "aot" else "compact tile of"
    render_exact(ffr, po, examples, "sierpinski_triangle_3d", [512, 512, 512], 200_000_000, 8192, 3,
                 ffr.JIT_OFF if kernel == "aot" else ffr.JIT_ON, expect)


def test_dense_512mib_compact_tile_overflow_bit_exact(ffr, po, examples, monkeypatch):
    """barnsley_fern at 8192^2 (512 MiB) with the tile capped at 32 MiB = 8192 rows of the ~40 000
    the fern touches: most rows take the DIRECT path next to rows that live in the tile."""
    monkeypatch.setenv("FFR_DIR_TILE_MB", "32")
    render_exact(ffr, po, examples, "barnsley_fern", [8192, 8192], 24_414 * 8192, 8192, 7, ffr.JIT_ON,
                 "compact tile of 8192 rows")
    # and across several launches of one context: rows keep their slots, the tile is zero between
    fl = ffr.Flame(examples.example_json("barnsley_fern", size=[8192, 8192]))
    r = ffr.BufferRenderer(fl, jit=ffr.JIT_ON)
    chains = 24_414
    for first, count in ((0, 5000), (5000, 12000), (17000, chains - 17000)):
        assert r.render_chains(first, count, 8192, base_seed=7)
    got = r.read_buffer()
    r.close()
    want, _ = oracle_cached(po, ffr, examples, "barnsley_fern", [8192, 8192], chains * 8192, 8192, 7)
    assert np.array_equal(got, want)


def coarse(counts, size, f):
    w, h = size
    return counts.reshape(h // f, f, w // f, f).astype(np.float64).sum(axis=(1, 3)).ravel()


def hist_l1(a, b):
    return float(np.abs(a / a.sum() - b / b.sum()).sum())


@pytest.mark.parametrize("name,size", [("csci6360_project", [4096, 4096]), ("tkoz_test3", [4096, 4096])])
def test_k1d_statistical_parity_at_config_size(ffr, po, examples, name, size):
    """cfg3 and the north-star target at their BASELINE buffer sizes, K1d (the kernel bench.py
    times) against the ORACLE: 2.0e8 samples each (24 414 chains x 8192). Class (iii) flames --
    libm vs CUDA transcendentals differ by ULPs and trajectories diverge chaotically -- so the
    check is statistical at equal sample count with a MEASURED tolerance: the noise floor is the
    largest distance between any two of three ORACLE runs with different seeds.
    Stated tolerances: L1 distance of the normalised 16x16-binned histograms (65 536 bins,
    ~3 000 samples each) <= 1.5 x floor; per-bin relative error on bins holding >= 0.01 % of the
    mass <= 1.5 x the oracle's own worst case + 1 %; s_plot and every xform selection count
    within 1.5 x the oracle's seed-to-seed spread + 5 sigma binomial; colour sums: mean colour
    per channel over the coarse bins within 1.5 x floor + 1e-3."""
    chains, L, f = 24_414, 8192, 16
    n = chains * L
    fl = ffr.Flame(examples.example_json(name, size=size))
    _, _, cells, cs = fl.layout()
    nthreads = min(32, os.cpu_count() or 8)

    def reduce_run(buf, st):
        c, col = ffr.split_counts_colors(buf, cells, cs - 1)
        cc = coarse(c, size, f)
        colc = None
        if col is not None:
            colc = np.stack([coarse(col[:, k].copy(), size, f) for k in range(cs - 1)], axis=1)
        return cc, colc, st

    r = ffr.BufferRenderer(fl, jit=ffr.JIT_ON)
    info = r.jit_info
    assert info["active"] and "K1d queue-scheduled kernel" in info["message"], info
    gruns = []
    # base seeds far apart: chain k of base seed s runs on splitmix64(s + k), so runs whose base
    # seeds differ by less than the chain count would share most of their chains
    for seed in (5_000_001, 9_000_001):
        r.clear()
        before = r.fetch_stats()
        assert r.render_chains(0, chains, L, base_seed=seed)
        st = r.stats
        buf = r.read_buffer()
        delta = {"s_plot": st["s_plot"] - before["s_plot"], "s_iter": st["s_iter"] - before["s_iter"],
                 "xf_dist": [a - b for a, b in zip(st["xf_dist"], before["xf_dist"])]}
        assert delta["s_iter"] == n
        c, _ = ffr.split_counts_colors(buf, cells, cs - 1)
        assert int(c.sum()) == delta["s_plot"]
        gruns.append(reduce_run(buf, delta))
        del buf, c
    r.close()
    oruns = []
    for seed in (1, 1_000_001, 2_000_001):
        o, st, _ = po.oracle_render(fl, chains, L, base_seed=seed, nthreads=nthreads)
        oruns.append(reduce_run(o, st))
        del o
    pairs = [(0, 1), (0, 2), (1, 2)]
    floor = max(hist_l1(oruns[i][0], oruns[j][0]) for i, j in pairs)
    worst = max(hist_l1(g[0], o[0]) for g in gruns for o in oruns)
    assert worst <= 1.5 * floor, (worst, floor)
    mean_o = sum(o[0] / o[0].sum() for o in oruns) / 3
    heavy = mean_o >= 1e-4
    assert heavy.sum() > 100

    def rel(a, b):
        return float((np.abs(a[heavy] / a.sum() - b[heavy] / b.sum()) / mean_o[heavy]).max())
    rfloor = max(rel(oruns[i][0], oruns[j][0]) for i, j in pairs)
    rworst = max(rel(g[0], o[0]) for g in gruns for o in oruns)
    assert rworst <= 1.5 * rfloor + 0.01, (rworst, rfloor)

    def spread(v):
        return max(v) - min(v)
    oplot = [o[2]["s_plot"] for o in oruns]
    p = np.mean(oplot) / n
    tol = 1.5 * spread(oplot) + 5 * (n * p * (1 - p)) ** 0.5 + 1
    for g in gruns:
        assert min(oplot) - tol <= g[2]["s_plot"] <= max(oplot) + tol
    for k in range(fl.desc.num_xform_ids):
        ox = [o[2]["xf_dist"][k] for o in oruns]
        q = np.mean(ox) / n
        tolk = 1.5 * spread(ox) + 5 * (n * q * (1 - q)) ** 0.5 + 1
        for g in gruns:
            assert min(ox) - tolk <= g[2]["xf_dist"][k] <= max(ox) + tolk, k
    if cs > 1:
        def cdist(a, b):
            m = (a[0] > 500) & (b[0] > 500)
            return float(np.abs(a[1][m] / a[0][m, None] - b[1][m] / b[0][m, None]).mean())
        cfloor = max(cdist(oruns[i], oruns[j]) for i, j in pairs)
        cworst = max(cdist(g, o) for g in gruns for o in oruns)
        assert cworst <= 1.5 * cfloor + 1e-3, (cworst, cfloor)


def _coarsen(counts, w, h, f):
    return counts.reshape(h // f, f, w // f, f).sum(axis=(1, 3))


@pytest.mark.parametrize("name,size,small", [("tkoz_test3", [4096, 4096], [512, 512]),
                                             ("csci6360_project", [8192, 8192], [512, 512]),
                                             ("csci6360_project", [4096, 4096], [256, 256])])
@pytest.mark.parametrize("kernel", ["aot", "k1d"])
def test_large_configs_size_independent_properties(ffr, examples, name, size, small, kernel):
    chains, L = 40_000, 2048
    jit = ffr.JIT_OFF if kernel == "aot" else ffr.JIT_ON
    fl = ffr.Flame(examples.example_json(name, size=size))
    _, _, cells, cs = fl.layout()
    r = ffr.BufferRenderer(fl, jit=jit)
    assert r.jit_info["active"] == (kernel == "k1d")
    assert r.render_chains(0, chains, L, base_seed=11)
    buf = r.read_buffer()
    st = r.stats
    s, m = r.histogram_sum_max()
    counts, colors = ffr.split_counts_colors(buf, cells, cs - 1)
    # checksum of checksums: histogram total == samples plotted == device reduction
    assert int(counts.sum()) == st["s_plot"] == s and int(counts.max()) == m
    assert st["s_iter"] == chains * L
    if colors is not None:
        # every colour coordinate is in [0,1], so 0 <= colour sum <= count in every cell
        assert (colors >= 0).all() and (colors <= counts[:, None] * (1 + 1e-12)).all()
        assert (colors[counts == 0] == 0).all()
    # the same chains at a small buffer: identical trajectories, so the coarsened full-size
    # histogram must equal the small one except for samples within rounding of a coarse cell
    # edge (index = trunc((x - lo) * size/(hi-lo) * (1-2^-52)) uses a different multiplier)
    fs = ffr.Flame(examples.example_json(name, size=small))
    rs = ffr.BufferRenderer(fs, jit=jit)
    assert rs.render_chains(0, chains, L, base_seed=11)
    sb = rs.read_buffer()
    sst = rs.stats
    rs.close()
    assert sst["s_plot"] == st["s_plot"] and sst["xf_dist"] == st["xf_dist"]
    assert sst["pt_min"] == st["pt_min"] and sst["pt_max"] == st["pt_max"]
    sc, _ = ffr.split_counts_colors(sb, small[0] * small[1], cs - 1)
    f = size[0] // small[0]
    coarse_ = _coarsen(counts, size[0], size[1], f).ravel()
    assert int(np.abs(coarse_.astype(np.int64) - sc.astype(np.int64)).sum()) <= 2e-6 * st["s_plot"] + 4
    # -i semantics at full size: two half renders added == one render (counts exact)
    r.clear()
    assert r.render_chains(0, chains // 2, L, base_seed=11)
    first = r.read_buffer().copy()
    r.clear()
    assert r.render_chains(chains // 2, chains - chains // 2, L, base_seed=11)
    r.add_buffer(first)
    both = r.read_buffer()
    r.close()
    c2, col2 = ffr.split_counts_colors(both, cells, cs - 1)
    assert np.array_equal(c2, counts)
    if colors is not None:
        np.testing.assert_allclose(col2, colors, rtol=1e-9, atol=1e-9)
