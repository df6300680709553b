"""The exact simplifier of the pure-affine kernel K1e (csrc/ffr_jit_host.cuh: generate_affine).

K1e's generated xform drops arithmetic the flame's coefficients make redundant (x*1, + 0*y,
the leading `0 +`, an identity post affine ...) -- but only where the result is provably the
same IEEE number, sign of zero included. This test does not trust the proof: it takes the
GENERATED `jaf_xform` text, compiles it as host C++ (gcc, x86-64 baseline: no FMA, like the
reference build) and compares it bit for bit with the oracle's XForm::applyIteration
(reference: types/xform.hpp:211-227, types/affine.hpp:104-110) on points that include +0, -0,
denormals, huge values and exact cancellations. No GPU needed (NVRTC only generates/compiles).
"""
import ctypes as C
import itertools
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import flames

WRAP = r"""
#include <cstring>
struct double2 { double x, y; };
struct float2 { float x, y; };
#define __device__
#define __forceinline__ inline
#define __constant__ static const
typedef %(jt)s JT;
typedef %(jt)s2 JPAIR;
#define JNX %(nx)d
#define JD %(d)d
%(consts)s
%(fn)s
extern "C" void apply(int k, const JT *pin, JT *pout, long n)
{
    for (long i = 0; i < n; ++i)
        jaf_xform((const JPAIR*)jtab + k, pin + i*JD, pout + i*JD);
}
"""


def build_host_xform(src, workdir):
    jt = "float" if "typedef float JT;" in src else "double"
    nx = int(re.search(r"#define JNX (\d+)", src).group(1))
    d = int(re.search(r"#define JD (\d+)", src).group(1))
    consts = "\n".join(re.findall(r"^__constant__ JT j(?:c|tab)\[.*$", src, re.M))
    fn = re.search(r"__device__ __forceinline__ void jaf_xform.*?\n}\n", src, re.S).group(0)
    path = os.path.join(workdir, "x.cpp")
    with open(path, "w") as f:
        f.write(WRAP % dict(jt=jt, nx=nx, d=d, consts=consts, fn=fn))
    so = os.path.join(workdir, "x.so")
    # no -march, no -ffast-math; contraction off for good measure (the baseline ISA has no FMA)
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, path],
                   check=True)
    return C.CDLL(so), nx, d, fn


def special_points(d, rng):
    vals = [0.0, -0.0, 1.0, -1.0, 0.5, -0.25, 5e-324, -5e-324, 1e-300, 3.0, -7.5, 1e19, -1e19,
            0.1, 1.0 / 3.0]
    pts = [list(p) for p in itertools.product(vals, repeat=d)] if d <= 2 else \
          [list(p) for p in itertools.product(vals[:9], repeat=d)]
    pts += rng.uniform(-2, 2, size=(4000, d)).tolist()
    pts += (rng.integers(-8, 9, size=(2000, d)) / 4.0).tolist()   # exact cancellations with the tables
    return np.array(pts, dtype=np.float64)


CASES = {
    "sierpinski-like (scale, identity post)": dict(pre="scale", post="identity"),
    "general pre, identity post": dict(pre="general", post="identity", nx=4),
    "general pre and post": dict(pre="general", post="general", nx=5),
    "identity pre, general post": dict(pre="identity", post="general"),
    "signed permutation pre": dict(pre="perm", post="scale"),
    "signed permutation post": dict(pre="general", post="perm"),
    "two variations": dict(pre="general", post="scale", weights=(0.75, 0.5)),
    "weight -1": dict(pre="general", post="identity", weights=(-1.0,)),
    "weights differ per xform": dict(pre="scale", post="general", weights=(0.8, 0.3), vary_weights=True),
    "-0.0 in a pre offset": dict(pre="scale", post="identity", neg_zero=("pre", 1, 1)),
    "-0.0 in a post offset": dict(pre="general", post="identity", neg_zero=("post", 0, 2)),
    "-0.0 offsets, weight 0.5": dict(pre="perm", post="perm", weights=(0.5,), neg_zero=("post", 1, 0)),
    "1-d": dict(dims=1, pre="general", post="scale"),
    "3-d no post": dict(dims=3, pre="scale", post="none", nx=4),
    "3-d no pre": dict(dims=3, pre="none", post="general", nx=4),
    "3-d neither, weight 0.5": dict(dims=3, pre="none", post="none", weights=(0.5,)),
    "3-d neither, weight 1": dict(dims=3, pre="none", post="none"),
    "3-d both": dict(dims=3, pre="general", post="perm", nx=8),
    "3-d no post, two variations": dict(dims=3, pre="perm", post="none", weights=(1.0, -1.0)),
}


@pytest.mark.parametrize("name", list(CASES), ids=lambda s: s.replace(" ", "_"))
@pytest.mark.parametrize("elem_size", [8, 4])
def test_generated_xform_is_bit_exact(ffr, po, name, elem_size):
    text = flames.affine_flame(**CASES[name])
    fl = ffr.Flame(text, elem_size=elem_size)
    try:
        src, _ = ffr.jit_compile(fl)
    except ffr.FfrError as e:
        if "libnvrtc not found" in str(e):
            pytest.skip("no NVRTC on this machine")
        raise
    assert "ffr_jit_affine.cuh" in src, "K1e was not chosen for a pure-affine flame"
    rng = np.random.default_rng(7)
    with tempfile.TemporaryDirectory() as wd:
        lib, nx, d, fn = build_host_xform(src, wd)
        pts = special_points(d, rng)
        if elem_size == 4:
            pts = pts.astype(np.float32).astype(np.float64)
        relies_on_no_negzero = "relies on: no coordinate of pin is -0.0" in src
        if relies_on_no_negzero:
            # the chain invariant the generator used: p = 2u-1 is never -0.0, and (checked
            # below) neither is any output of the generated function
            pts = pts[~((pts == 0) & np.signbit(pts)).any(axis=1)]
        seeds = np.zeros(len(pts), dtype=np.uint64)
        for k in range(nx):
            if elem_size == 8:
                want = po.oracle_iterate_points(fl, k, seeds, pts)
                got = np.empty_like(pts)
                lib.apply(k, pts.ctypes.data_as(C.c_void_p), got.ctypes.data_as(C.c_void_p), C.c_long(len(pts)))
                a, b = got.view(np.uint64), want.view(np.uint64)
            else:
                if not po.have_ref(4):
                    pytest.skip("float reference build not present")
                want = po.ref_iterate_points(text, k, seeds, pts, elem_size=4).astype(np.float32)
                p32 = pts.astype(np.float32)
                got = np.empty_like(p32)
                lib.apply(k, p32.ctypes.data_as(C.c_void_p), got.ctypes.data_as(C.c_void_p), C.c_long(len(pts)))
                a, b = got.view(np.uint32), want.view(np.uint32)
            if relies_on_no_negzero:
                assert not ((got == 0) & np.signbit(got)).any(), "the invariant does not hold"
            bad = np.nonzero((a != b).any(axis=1))[0]
            assert len(bad) == 0, (name, k, pts[bad[:3]], got[bad[:3]], want[bad[:3]], fn)


def test_simplifier_output_shapes(ffr):
    """What the rules are expected to leave (documentation by example)."""
    def xform_of(**kw):
        src, _ = ffr.jit_compile(ffr.Flame(flames.affine_flame(**kw)))
        return re.search(r"void jaf_xform.*?\n}\n", src, re.S).group(0), src
    try:
        fn, src = xform_of(pre="scale", post="identity")
    except ffr.FfrError as e:
        if "libnvrtc not found" in str(e):
            pytest.skip("no NVRTC on this machine")
        raise
    # sierpinski shape: one multiply and one add per coordinate, nothing for the variation sum
    # or the identity post affine
    assert fn.count("*x") == 2 and "const T v0 = t0;" in fn and "pout[0] = v0;" in fn
    assert "#define JNPAIR 1" in src
    # a -0.0 offset keeps that row in the reference's full form
    fn, _ = xform_of(pre="scale", post="identity", neg_zero=("pre", 1, 1))
    assert "(((T)0.0 + " in fn
    # flames K1e does not cover fall back to the general generator
    src, _ = ffr.jit_compile(ffr.Flame(flames.divergent_flame()))
    assert "ffr_jit_affine.cuh" not in src


def test_flames_k1e_refuses(ffr):
    """Outside K1e's proof obligations the general generator is used (same results, slower):
    coefficients too large to prove every intermediate finite through the unchecked settle
    iterations, bounds beyond the bad value threshold (in-bounds must imply not-bad), xforms of
    different shape, colours, a final xform."""
    import json

    def kernel_of(fl):
        try:
            src, _ = ffr.jit_compile(ffr.Flame(json.dumps(fl)))
        except ffr.FfrError as e:
            if "libnvrtc not found" in str(e):
                pytest.skip("no NVRTC on this machine")
            raise
        return "K1e" if "ffr_jit_affine.cuh" in src else "general"

    base = json.loads(flames.affine_flame(pre="general", post="identity"))
    assert kernel_of(base) == "K1e"
    big = json.loads(json.dumps(base))
    big["xforms"][0]["pre_affine"]["A"][0][0] = 3.0e6
    assert kernel_of(big) == "general"
    wide = json.loads(json.dumps(base))
    wide["bounds"][0] = [-1.0e21, 1.0]      # the flame model already refuses such bounds
    with pytest.raises(ffr.FfrError):
        kernel_of(wide)
    ragged = json.loads(json.dumps(base))
    ragged["xforms"][1]["variations"].append({"name": "linear", "weight": 0.25})
    assert kernel_of(ragged) == "general"
    col = json.loads(json.dumps(base))
    col["color_dimensions"] = 1
    col["xforms"][0]["color"] = [0.5]
    assert kernel_of(col) == "general"
    fin = json.loads(json.dumps(base))
    fin["final_xform"] = {"variations": [{"name": "linear", "weight": 1.0}]}
    assert kernel_of(fin) == "general"
    # mildly expanding maps are fine: alpha^55 stays far below DBL_MAX
    mild = json.loads(json.dumps(base))
    mild["xforms"][0]["pre_affine"]["A"][0][0] = 50.0
    assert kernel_of(mild) == "K1e"
