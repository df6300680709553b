"""include/ffr_buffer_renderer.hpp: the reference's BufferRenderer<dims> surface
(src/renderers/buffer_renderer.hpp:252-583) as a C++ class over the C ABI, driven by a C++ host
(tests/cpp/buffer_renderer_host.cpp) written the way src/ffr_buf.cpp:146-274 uses the reference's
class. CPU part: the header compiles, constructor errors are the reference's. GPU part: the
buffer file and every getter against the oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "flame-fractal-renderer_b200")


@pytest.fixture(scope="module")
def host_exe(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("cpp") / "buffer_renderer_host")
    cxx = os.environ.get("CXX", "g++")
    p = subprocess.run([cxx, "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "cpp", "buffer_renderer_host.cpp"), "-o", exe,
                        "-L", PKG, "-lffr_cuda", "-Wl,-rpath," + PKG], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    return exe


def _construct(exe, tmp_path, text):
    f = tmp_path / "f.json"
    f.write_text(text)
    p = subprocess.run([exe, "--construct", str(f)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    return json.loads(p.stdout)


def test_constructor_errors_are_the_references(host_exe, examples, tmp_path):
    """Flame validation failures surface as JsonError with the reference's texts
    (types/flame.hpp:91-210), before any device is touched."""
    r = _construct(host_exe, tmp_path, "not json")
    assert r == {"constructed": False, "kind": "JsonError", "what": r["what"]} and "parse" in r["what"]
    bad = json.loads(examples.example_json("sierpinski_triangle"))
    bad["bounds"] = bad["bounds"][:1]
    r = _construct(host_exe, tmp_path, json.dumps(bad))
    assert (r["kind"], r["what"]) == ("JsonError", "Flame(): incorrect bounds length")
    big = json.loads(examples.example_json("sierpinski_triangle"))
    big["size"] = [65536, 16]          # constants.hpp:33 max_dim = 65535
    r = _construct(host_exe, tmp_path, json.dumps(big))
    assert r["kind"] == "JsonError" and r["what"].endswith("out of range")


def test_no_device_no_fallback(host_exe, examples, tmp_path):
    """Without a GPU the constructor throws: there is no CPU path behind the class."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    r = _construct(host_exe, tmp_path, examples.example_json("sierpinski_triangle", size=[64, 64]))
    assert r["constructed"] is False and r["kind"] == "runtime_error" and "no CPU fallback" in r["what"]


def _run(exe, flame_path, out, samples, batch, seed, split=0, inputs=()):
    p = subprocess.run([exe, str(flame_path), str(out), str(samples), str(batch), str(seed), str(split)]
                       + [str(i) for i in inputs], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    return json.loads(p.stdout)


@pytest.mark.gpu
@pytest.mark.parametrize("name,size", [("sierpinski_triangle", [200, 150]), ("sierpinski_triangle_3d", [40, 48, 56]),
                                       ("flam3_test_1", [160, 120])])
def test_render_write_and_getters_match_the_oracle(host_exe, ffr, po, examples, tmp_path, name, size):
    """render() + writeBuffer() + the statistics getters, bit for bit (pure-affine and IEEE-only
    flames, SURVEY Q6 classes i/ii)."""
    text = examples.example_json(name, size=size)
    fpath = tmp_path / "flame.json"
    fpath.write_text(text)
    out = tmp_path / "a.buf"
    samples, batch, seed = 1_234_567, 4096, 77
    got = _run(host_exe, fpath, out, samples, batch, seed)
    fl = ffr.Flame(text)
    want, st, ok = po.oracle_render_samples(fl, samples, batch, base_seed=seed, nthreads=8)
    a = np.fromfile(out, dtype=np.uint64)
    assert got["ok"] and ok and np.array_equal(a, want)
    chains = (samples + batch - 1) // batch
    assert got["batches"] == chains and got["workers"] == 1 and got["next_seed"] == seed + chains
    assert (got["s_iter"], got["s_plot"]) == (st["s_iter"], st["s_plot"]) == (samples, got["sum"])
    assert got["xf_dist"] == [int(x) for x in st["xf_dist"]]
    cells = int(np.prod(size))
    counts = a.reshape(cells, -1)[:, 0]
    assert (got["min"], got["max"], got["cell_sum"]) == (int(counts.min()), int(counts.max()), int(counts.sum()))
    assert got["twice_sum"] == 2 * got["sum"]
    assert (got["cells"], got["cell_size"], got["dims"], got["color_dims"]) == (cells, 1, len(size), 0)
    assert got["bad"] == 0 and got["bad_pts"] == 0
    for d in range(len(size)):
        assert got["extremes"][d] == [st["pt_min"][d], st["pt_max"][d]]
    md, mi, _, _ = fl.layout()
    assert got["mult_d"] == [float(x) for x in md[:len(size)]] and got["mult_i"] == [int(x) for x in mi[:len(size)]]
    # argument checks of render(): the reference's messages (buffer_renderer.hpp:281-289)
    assert got["e_threads0"] == "BufferRenderer::render(): threads must be positive"
    assert got["e_threads"] == "BufferRenderer::render(): too many threads"
    assert got["e_batch"] == "BufferRenderer::render(): batch size too small"
    assert got["e_vec"] == "BufferRenderer::addBuffer(): sizes do not match"


@pytest.mark.gpu
def test_two_calls_continue_the_chain_numbering_and_inputs_add(host_exe, ffr, po, examples, tmp_path):
    """render() then renderSeeded() on one object never repeat a chain (the second call's first
    batch is seed + batches of the first), and addBuffer(istream) adds input files like -i."""
    text = examples.example_json("barnsley_fern", size=[128, 96])
    fpath = tmp_path / "flame.json"
    fpath.write_text(text)
    fl = ffr.Flame(text)
    first, rest, batch, seed = 300_000, 500_001, 1000, 5
    out = tmp_path / "a.buf"
    got = _run(host_exe, fpath, out, first + rest, batch, seed, split=first)
    w1, _, _ = po.oracle_render_samples(fl, first, batch, base_seed=seed)
    w2, _, _ = po.oracle_render_samples(fl, rest, batch, base_seed=seed + first // batch)
    a = np.fromfile(out, dtype=np.uint64)
    assert np.array_equal(a, w1 + w2)
    assert got["batches"] == first // batch + (rest + batch - 1) // batch and got["s_iter"] == first + rest
    out2 = tmp_path / "b.buf"
    got = _run(host_exe, fpath, out2, 0, batch, seed, inputs=[out, out])
    assert got["added"] and got["batches"] == 0
    assert np.array_equal(np.fromfile(out2, dtype=np.uint64), 2 * a)
    short = tmp_path / "short.buf"
    short.write_bytes(a.tobytes()[:-8])
    got = _run(host_exe, fpath, tmp_path / "c.buf", 0, batch, seed, inputs=[short])
    assert got["added"] is False and got["sum"] == 0      # short read: false, buffer untouched


@pytest.mark.gpu
def test_colour_buffer_layout_through_the_class(host_exe, ffr, po, examples, tmp_path):
    """A flame with colour dimensions: cells x [count, c0, c1] (buffer_renderer.hpp:60-64), counts
    bit for bit, colour sums to the order of the additions (the reference's multithreaded sums are
    order dependent too): relative 1e-12."""
    fl_json = json.loads(examples.example_json("sierpinski_triangle", size=[96, 64]))
    fl_json["color_dimensions"] = 2
    fl_json["color_speed"] = 0.25
    for k, xf in enumerate(fl_json["xforms"]):
        xf["color"] = [k / 2.0, 1.0 - k / 4.0]
    text = json.dumps(fl_json)
    fpath = tmp_path / "flame.json"
    fpath.write_text(text)
    out = tmp_path / "a.buf"
    samples, batch, seed = 400_000, 2048, 11
    got = _run(host_exe, fpath, out, samples, batch, seed)
    fl = ffr.Flame(text)
    want, st, _ = po.oracle_render_samples(fl, samples, batch, base_seed=seed)
    cells = 96 * 64
    assert (got["cells"], got["cell_size"], got["color_dims"]) == (cells, 3, 2)
    a = np.fromfile(out, dtype=np.uint64).reshape(cells, 3)
    w = want.reshape(cells, 3)
    assert np.array_equal(a[:, 0], w[:, 0]) and got["sum"] == int(w[:, 0].sum()) == st["s_plot"]
    assert np.allclose(a[:, 1:].copy().view(np.float64), w[:, 1:].copy().view(np.float64), rtol=1e-12, atol=0.0)
