"""Host logic + C ABI surface, no GPU: the library loads (cudart is linked statically),
exports every symbol include/*.h declares, and the flame model behaves like the reference's
constructors (types/flame.hpp:91-210, types/xform.hpp:71-172, variations.hpp factories)."""
import ctypes
import json
import os
import re
import subprocess

import pytest

import flames

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    syms = []
    for h in ("ffr_cuda.h", "ffr_flame.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        syms += re.findall(r"\b(ffr_[a-z0-9_]+)\s*\(", text)
    return sorted(set(s for s in syms if s not in ("ffr_progress_cb",)))


def test_library_exports_every_declared_symbol(ffr):
    lib = ctypes.CDLL(ffr.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), s
    # and the binding table covers exactly the declared ABI
    assert sorted(ffr.ABI) == syms


def test_version_and_no_gpu_behaviour(ffr, examples):
    L = ffr.lib()
    assert b"sm_100a" in L.ffr_cuda_version()
    n = L.ffr_cuda_device_count()
    assert n >= 0
    if n == 0:
        fl = ffr.Flame(examples.example_json("sierpinski_triangle"))
        with pytest.raises(ffr.FfrError, match="no CPU fallback"):
            ffr.BufferRenderer(fl)


def test_product_does_not_touch_the_oracle():
    """The product path must never import/link/execute anything under oracle/."""
    pkg = os.path.join(ROOT, "flame-fractal-renderer_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cpp", ".cu", ".cuh", ".hpp", ".h")) or fn == "Makefile":
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "pyoracle" not in text and "ffr_oracle" not in text and "libffr_ref" not in text, fn
    out = subprocess.run(["ldd", os.path.join(pkg, "libffr_cuda.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_chain_seed_is_splitmix64(ffr, po):
    L = ffr.lib()
    for base, k in ((1, 0), (1, 5), (2**64 - 1, 3), (123456789, 10**12)):
        assert L.ffr_chain_seed(base, k) == po.oracle().oracle_splitmix64((base + k) % 2**64)


def test_reference_batch_heuristic(ffr):
    L = ffr.lib()
    assert L.ffr_reference_batch_size(0) == 4096
    assert L.ffr_reference_batch_size(10**8) == 390625
    assert L.ffr_reference_batch_size(10**11) == 1 << 20


def test_xform_order_and_weights(ffr, examples):
    # SURVEY.md section 4 pins
    f = ffr.Flame(examples.example_json("barnsley_fern"))
    assert f.xform_ids == [1, 2, 3, 0]
    assert [repr(x) for x in f.cumulative_weights] == [
        "0.85", "0.9199999999999999", "0.99", "1.0"]
    f = ffr.Flame(examples.example_json("csci6360_project"))
    assert f.xform_ids == [4, 1, 0, 3, 2]
    assert [repr(x) for x in f.cumulative_weights] == [
        "0.32592592592592595", "0.6222222222222222", "0.7703703703703704", "0.9111111111111111", "1.0"]
    f = ffr.Flame(examples.example_json("sierpinski_triangle"))
    assert repr(f.layout()[0][0]) == "511.9999999999999"
    f = ffr.Flame(examples.example_json("barnsley_fern"))
    assert repr(f.layout()[0][0]) == "51.19999999999999"


def test_size_override(ffr, examples):
    text = examples.example_json("csci6360_project")
    f = ffr.Flame(text, size=[4096, 4096])
    assert f.size == [4096, 4096]
    assert f.layout()[2] == 4096 * 4096
    with pytest.raises(ffr.FfrError):
        ffr.Flame(text, size=[4096])


def test_comments_and_number_forms(ffr):
    text = """
    { // line comment
      "dimensions": 2, /* block
      comment */ "size": [16, 32.0], "bounds": [[-1, 1], [0, 2.5e0]],
      "xforms": [ {"weight": 1, "variations": [{"name": "linear", "weight": 1e0}],
                   "pre_affine": {"A": [[5e-1, 0], [0, -0.5]], "b": [0, 1]}} ] }
    """
    f = ffr.Flame(text)
    assert f.size == [16, 32]
    x = f.desc.xforms[0]
    assert list(x.pre_A)[:4] == [0.5, 0.0, 0.0, -0.5] and x.has_pre and not x.has_post
    assert list(x.post_A)[:4] == [1.0, 0.0, 0.0, 1.0]


def base_flame():
    return json.loads(flames.variation_flame("linear", dims=2, color=False))


@pytest.mark.parametrize("mutate,msg", [
    (lambda f: f.update(size=[0, 4]), "size[0] out of range"),
    (lambda f: f.update(size=[4, 70000]), "size[1] out of range"),
    (lambda f: f.update(size=[4]), "incorrect size length"),
    (lambda f: f.update(bounds=[[1, 1], [0, 1]]), "low >= high"),
    (lambda f: f.update(bounds=[[-1e11, 1], [0, 1]]), "out of range"),
    (lambda f: f.update(xforms=[]), "no xforms"),
    (lambda f: f.update(dimensions=4), "not supported"),
    (lambda f: f.update(color_dimensions=128), "too many color dimensions"),
    (lambda f: f.update(color_dimensions=3.0), "not int"),
    (lambda f: f.update(color_speed=1.5), "color speed out of range"),
    (lambda f: f["xforms"][0].update(weight=-1), "weight is negative"),
    (lambda f: f["xforms"][0]["variations"][0].update(name="nope"), "unknown variation"),
    (lambda f: [x.update(weight=0) for x in f["xforms"]], "no xforms remaining"),
    (lambda f: f["xforms"][0].update(pre_affine={"A": [[1, 0]]}), "A is wrong size"),
])
def test_validation_errors(ffr, mutate, msg):
    f = base_flame()
    mutate(f)
    with pytest.raises(ffr.FfrError, match=re.escape(msg)):
        ffr.Flame(json.dumps(f))


def test_variation_parameter_validation(ffr):
    f = base_flame()
    f["xforms"][0]["variations"][0] = {"name": "boarders", "weight": 1, "prob": 1.5}
    with pytest.raises(ffr.FfrError, match="boarders probability"):
        ffr.Flame(json.dumps(f))
    f["xforms"][0]["variations"][0] = {"name": "spherical_p", "weight": 1, "norm": 0}
    with pytest.raises(ffr.FfrError, match="norm <= 0"):
        ffr.Flame(json.dumps(f))
    f["xforms"][0]["variations"][0] = {"name": "fisheye", "weight": 1}
    with pytest.raises(ffr.FfrError, match="key does not exist: addval"):
        ffr.Flame(json.dumps(f))
    # 2-d variation in 3-d needs distinct in-range axes (variations.hpp:72-88)
    g = json.loads(flames.variation_flame("swirl", dims=3, color=False))
    g["xforms"][1]["variations"][0]["axis_y"] = g["xforms"][1]["variations"][0]["axis_x"]
    with pytest.raises(ffr.FfrError, match="axes are not distinct"):
        ffr.Flame(json.dumps(g))
    h = json.loads(flames.variation_flame("linear", dims=1, color=False))
    h["xforms"][0]["variations"][0] = {"name": "swirl", "weight": 1}
    with pytest.raises(ffr.FfrError, match="unknown variation"):
        ffr.Flame(json.dumps(h))


def test_zero_weight_variations_and_xforms_dropped(ffr):
    f = base_flame()
    f["xforms"][1]["variations"].append({"name": "bubble", "weight": 0.0, "addval": 4})
    f["xforms"][2]["weight"] = 0.0
    fl = ffr.Flame(json.dumps(f))
    assert fl.desc.num_xforms == 2 and fl.desc.num_xform_ids == 3
    assert fl.xform_ids == [0, 1]
    assert fl.desc.xforms[1].num_vars == 2


def test_histogram_size_guard(ffr):
    # cells >= 2^48 throws in the reference (buffer_renderer.hpp:132-133); with dims <= 3 and
    # sizes <= 65535 the largest buffer, 65535^3 cells, is just below it and must be accepted
    f = json.loads(flames.variation_flame("linear", dims=3, color=False))
    f["size"] = [65535, 65535, 65535]
    fl = ffr.Flame(json.dumps(f))
    assert fl.layout()[2] == 65535 ** 3 < 2 ** 48


def test_variation_name_table(ffr):
    L = ffr.lib()
    for i, name in enumerate(flames.ALL_VARIATIONS):
        op = L.ffr_var_op_from_name(name.encode())
        assert 1 <= op <= 98
        assert L.ffr_var_name(op) == name.encode()
    assert L.ffr_var_op_from_name(b"nope") == 0
