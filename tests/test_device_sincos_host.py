"""sincos_core (csrc/ffr_device.cuh), the double precision sin/cos every render kernel uses for
the reference's sin / cos / math::sincosg calls (utils/math.hpp:21-24, variations.hpp), checked
without a GPU: the function's TEXT is taken out of the header and compiled as host C++. It is
written with explicit fma() only (correctly rounded in glibc as on the device, x86-64 baseline
build: no contraction of anything else), so the host build computes the very bits the device
computes. Tolerance stated by the kernel: <= 2 ULP of glibc's result on the fast path
|x| <= 105615 (the GPU twin of this test is test_gpu_parity.py::test_device_sin_cos_accuracy).
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
HDR = os.path.join(HERE, "..", "flame-fractal-renderer_b200", "csrc", "ffr_device.cuh")

WRAP = r"""
#include <cmath>
#include <cstring>
#include <cstdint>
#define __device__
#define __forceinline__ inline
#define __constant__ static const
static inline int __double2loint(double x) { uint64_t u; std::memcpy(&u,&x,8); return (int)(uint32_t)u; }
static inline int __double2hiint(double x) { uint64_t u; std::memcpy(&u,&x,8); return (int)(uint32_t)(u >> 32); }
static inline double __hiloint2double(int hi, int lo)
{
    uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double x; std::memcpy(&x,&u,8); return x;
}
using std::fma;
%(consts)s
%(fn)s
extern "C" void run(const double *x, double *s, double *c, long n)
{
    for (long i = 0; i < n; ++i)
        sincos_core(x[i],s[i],c[i]);
}
extern "C" double fast_bound(void) { return FFR_SC[19]; }
"""


@pytest.fixture(scope="module")
def host_sincos(tmp_path_factory):
    text = open(HDR).read()
    consts = re.search(r"__constant__ double FFR_SC\[20\] = \{.*?\};", text, re.S).group(0)
    fn = re.search(r"__device__ __forceinline__ void sincos_core\(double x.*?\n}\n", text, re.S).group(0)
    d = tmp_path_factory.mktemp("sincos")
    src = d / "sc.cpp"
    src.write_text(WRAP % dict(consts=consts, fn=fn))
    so = d / "sc.so"
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", str(so), str(src)],
                   check=True)
    lib = C.CDLL(str(so))
    lib.run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
    lib.fast_bound.restype = C.c_double

    def run(x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        s = np.empty_like(x)
        c = np.empty_like(x)
        lib.run(x.ctypes.data, s.ctypes.data, c.ctypes.data, len(x))
        return s, c
    return run, lib.fast_bound()


def _ulps(a, b):
    return np.abs(a - b) / np.spacing(np.abs(b))


def test_fast_path_within_two_ulp_of_glibc(host_sincos):
    run, bound = host_sincos
    assert bound == 105615.0
    rng = np.random.default_rng(11)
    xs = np.concatenate([
        rng.uniform(-4, 4, 400000), rng.uniform(-1e3, 1e3, 400000), rng.uniform(-bound, bound, 400000),
        rng.uniform(-1, 1, 100000) * 10.0 ** rng.uniform(-300, 0, 100000),
        # near the multiples of pi/2, where the reduction decides everything
        np.pi / 2 * np.arange(-67000, 67000) + rng.uniform(-1e-9, 1e-9, 134000),
        np.pi / 2 * rng.integers(-67000, 67000, 100000) * (1 + rng.uniform(-1e-15, 1e-15, 100000)),
        np.array([bound, -bound, np.pi / 4, -np.pi / 4, np.nextafter(np.pi / 4, 1), 1e-300, 5e-324])])
    xs = xs[np.abs(xs) <= bound]
    s, c = run(xs)
    ws, wc = np.sin(xs), np.cos(xs)
    assert _ulps(s, ws).max() <= 2.0
    assert _ulps(c, wc).max() <= 2.0
    # and nowhere near the bound on the bulk: the mean error stays a fraction of an ULP
    assert _ulps(s, ws).mean() < 0.4 and _ulps(c, wc).mean() < 0.4


def test_signed_zero_and_exact_values(host_sincos):
    run, _ = host_sincos
    s, c = run(np.array([0.0, -0.0]))
    assert s[0] == 0.0 and not np.signbit(s[0])
    assert s[1] == 0.0 and np.signbit(s[1])          # sin(-0) = -0 like glibc
    assert c[0] == 1.0 and c[1] == 1.0


def test_symmetry(host_sincos):
    """sin is odd and cos is even bit for bit (the reduction rounds to nearest, ties cannot occur
    for an irrational period)."""
    run, bound = host_sincos
    rng = np.random.default_rng(5)
    xs = rng.uniform(0, bound, 200000)
    s0, c0 = run(xs)
    s1, c1 = run(-xs)
    assert np.array_equal(s0, -s1)
    assert np.array_equal(c0, c1)
