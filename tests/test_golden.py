"""The oracle restatement (oracle/ffr_oracle.c) + host flame model against the committed
golden vectors, which are outputs of the UNMODIFIED reference (tests/golden/make_golden.py).
Runs anywhere: needs neither /root/reference nor a GPU."""
import hashlib
import json
import os

import numpy as np
import pytest

import flames

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "golden.json")) as f:
    GOLD = json.load(f)
P = GOLD["params"]


def clean(st):
    def f(x):
        if isinstance(x, float):
            return repr(x)
        if isinstance(x, list):
            return [f(y) for y in x]
        return x
    return {k: f(v) for k, v in st.items() if k not in ("bad_xf", "bad_pt")}


def check_case(ffr, po, text, gold):
    fl = ffr.Flame(text)
    assert fl.xform_ids == gold["ids"]
    assert [repr(x) for x in fl.cumulative_weights] == gold["cw"]
    md, mi, cells, cs = fl.layout()
    assert [repr(x) for x in md] == gold["mult_d"]
    assert mi == gold["mult_i"] and cells == gold["cells"] and cs == gold["cell_size"]
    po.set_nan_emulation(True)  # the reference binary's NaN behaviour, see ffr_oracle.c
    try:
        buf, st, ok = po.oracle_render(fl, P["chains"], P["chain_len"], base_seed=P["base_seed"],
                                       last_len=P["last_len"], bv_limit=1 << 20)
    finally:
        po.set_nan_emulation(False)
    assert ok == gold["ok"]
    assert clean(st) == gold["stats"]
    assert hashlib.sha256(buf.tobytes()).hexdigest() == gold["sha256"]
    assert int(buf.reshape(-1, cs)[:, 0].sum()) == gold["hist_sum"]


def test_isaac_known_answers(po):
    for seed, words in GOLD["isaac"].items():
        got = ["%016x" % int(x) for x in po.oracle_isaac_words(int(seed), len(words))]
        assert got == words
    # SURVEY.md section 4 pins
    assert GOLD["isaac"]["1"][:4] == ["3dc7e2e12622c959", "262ccb29475eb0cd",
                                      "62eb77756c571e1a", "326be2ff22a85a27"]
    assert GOLD["isaac"]["2"][:4] == ["1bd98216b66da880", "7b5eb80a1b3386e0",
                                      "b394f42a1b0e0eda", "97941dc236c3db23"]


def test_survey_md5_pins_recorded():
    # rng::setSeed(12345 / 999) + renderSeeded(1e7, 1<<20, 256) on sierpinski 512^2
    assert GOLD["raw_pins"]["12345"] == {"md5": "1d00f9b7761966e8c7244d875f2c0829", "max": 595}
    assert GOLD["raw_pins"]["999"] == {"md5": "343c317bfdd03a70fac15e645aa029b4", "max": 605}


@pytest.mark.parametrize("name", sorted(GOLD["examples"]))
def test_examples_match_reference(ffr, po, examples, name):
    size = [48, 48, 48] if name.endswith("3d") else None
    check_case(ffr, po, examples.example_json(name, size=size), GOLD["examples"][name])


@pytest.mark.parametrize("key", sorted(k for k in GOLD["variations"] if "/" in k))
def test_variations_match_reference(ffr, po, key):
    name, d = key.split("/")
    d = int(d)
    check_case(ffr, po, flames.variation_flame(name, dims=d, final=(d == 3)), GOLD["variations"][key])


def test_edge_flames_match_reference(ffr, po):
    check_case(ffr, po, flames.divergent_flame(), GOLD["variations"]["divergent"])
    check_case(ffr, po, flames.one_d_flame(), GOLD["variations"]["one_d"])
    check_case(ffr, po, flames.many_xforms_flame(), GOLD["variations"]["many_xforms"])


def test_golden_covers_every_variation():
    names = {k.split("/")[0] for k in GOLD["variations"] if "/" in k}
    assert names == set(flames.ALL_VARIATIONS)
    assert len(names) == 98


def test_threaded_oracle_counts_are_exact(ffr, po, examples):
    fl = ffr.Flame(examples.example_json("tkoz_test3", size=[96, 54]))
    a, sa, _ = po.oracle_render(fl, 64, 700, base_seed=3, nthreads=1)
    b, sb, _ = po.oracle_render(fl, 64, 700, base_seed=3, nthreads=8)
    _, _, cells, cs = fl.layout()
    ca, cola = ffr.split_counts_colors(a, cells, cs - 1)
    cb, colb = ffr.split_counts_colors(b, cells, cs - 1)
    assert np.array_equal(ca, cb)
    np.testing.assert_allclose(cola, colb, rtol=1e-12, atol=1e-12)
    for k in ("s_iter", "s_plot", "xf_dist", "pt_min", "pt_max"):
        assert sa[k] == sb[k]


with open(os.path.join(HERE, "golden", "golden_img.json")) as f:
    GOLD_IMG = json.load(f)


@pytest.mark.parametrize("name", sorted(GOLD_IMG["tonemap"]))
def test_oracle_tonemap_matches_reference_pixels(ffr, po, examples, name):
    """f1 pin: oracle_tonemap against the pixels the reference's own render_image()
    (ffr_img.cpp:199-309 over image_renderer.hpp:112-192, compiled by `make -C oracle refimg`)
    produced for the same seeded buffer -- every mode x bit depth x gamma, bit for bit."""
    gold = GOLD_IMG["tonemap"][name]
    pp = GOLD_IMG["params"]
    fl = ffr.Flame(examples.example_json(name, size=gold["size"]))
    po.set_nan_emulation(True)
    try:
        raw, _, _ = po.oracle_render(fl, pp["chains"], pp["chain_len"], base_seed=pp["base_seed"])
    finally:
        po.set_nan_emulation(False)
    assert hashlib.sha256(raw.tobytes()).hexdigest() == gold["buffer_sha256"]
    w, h = gold["size"]
    for key, want in gold["cases"].items():
        mode, bits, gamma = key.split("/")
        img, info = po.oracle_tonemap(raw, w, h, fl.color_dims, int(mode), bits=int(bits),
                                      gamma=float(gamma))
        assert hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest() == want["sha256"], key
        assert info["hist_min"] == want["hist_min"] and info["hist_max"] == want["hist_max"]


def test_flame_echo_matches_reference(ffr):
    """The `flame: ...` line of ffr_buf.cpp:129 (nlohmann's compact dump through
    utils/json.cpp:203-207): ffr_flame_json_echo against the text the reference printed."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden_img
    texts = make_golden_img.echo_texts()
    assert sorted(texts) == sorted(GOLD_IMG["echo"])
    for key, text in texts.items():
        assert ffr.flame_json_echo(text) == GOLD_IMG["echo"][key], key


def test_oracle_tonemap_properties(ffr, po, examples):
    """ffr-img pixel math (oracle restatement, pinned above): exact identities
    that follow from src/ffr_img.cpp:236-243 -- the brightest cell is the top code, empty cells
    are 0, gamma 1 grey equals floor(log(1+n)/log(1+max) * 256(1-2^-52)), output is monotone
    in the count."""
    fl = ffr.Flame(examples.example_json("barnsley_fern", size=[96, 64]))
    raw, st, _ = po.oracle_render(fl, 40, 2000, base_seed=2)
    counts = raw.reshape(64, 96)
    for bits, top in ((8, 255), (16, 65535)):
        img, info = po.oracle_tonemap(raw, 96, 64, 0, 2, bits=bits, gamma=1.0)
        assert info["hist_max"] == counts.max() and info["hist_min"] == counts.min()
        assert img[counts == counts.max()].min() == top
        assert (img[counts == 0] == 0).all()
        scale = (top + 1) * (1.0 - 2.0 ** -52)
        want = np.floor(np.log(1.0 + counts.astype(np.float64)) / info["scaler_max"] * scale)
        assert np.array_equal(img, np.minimum(want, top).astype(img.dtype))
        order = np.argsort(counts.ravel(), kind="stable")
        assert (np.diff(img.ravel()[order].astype(np.int64)) >= 0).all()
    mono, _ = po.oracle_tonemap(raw, 96, 64, 0, 1)
    assert np.array_equal(mono, np.where(counts != 0, 255, 0).astype(np.uint8))
