"""oracle/pyoracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes access to the two CPU checkers:
  libffr_oracle.so        our plain C restatement (oracle/ffr_oracle.c)
  _ref/libffr_ref.so      the unmodified reference behind oracle/ref_harness.cpp
  _ref/libffr_refimg.so   the reference's tone map (image_renderer.hpp + render_image() of
                          ffr_img.cpp) and its `os << Json` echo, behind ref_img_harness.cpp
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm import
this module. Nothing here touches /root/reference at run time: _ref/libffr_ref.so is
prebuilt by `make -C oracle ref` in the build container and travels with the repo.
"""

import ctypes as C
import importlib
import json
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(_HERE, "libffr_oracle.so")
REF_LIB = os.path.join(_HERE, "_ref", "libffr_ref.so")
REF_LIB_F32 = os.path.join(_HERE, "_ref", "libffr_ref_f32.so")  # float/uint32_t configuration
REFIMG_LIB = os.path.join(_HERE, "_ref", "libffr_refimg.so")    # tone map + "flame:" echo (ref_img_harness.cpp)

ffr = importlib.import_module("flame-fractal-renderer_b200")

_u64p = C.POINTER(C.c_uint64)
_f64p = C.POINTER(C.c_double)
_oracle = None
_ref = {}


def have_ref(elem_size=8):
    return os.path.exists(REF_LIB if elem_size == 8 else REF_LIB_F32)


def oracle():
    global _oracle
    if _oracle is None:
        l = C.CDLL(ORACLE_LIB)
        l.oracle_splitmix64.restype = C.c_uint64
        l.oracle_splitmix64.argtypes = [C.c_uint64]
        l.oracle_isaac_words.argtypes = [C.c_uint64, C.c_uint64, _u64p]
        l.oracle_rand_nums.argtypes = [C.c_uint64, C.c_uint64, _f64p]
        l.oracle_render_chains.restype = C.c_int
        l.oracle_render_chains.argtypes = [C.POINTER(ffr.FfrFlameDesc)] + [C.c_uint64] * 6 + [
            C.c_void_p, C.POINTER(ffr.FfrStats), C.c_int]
        l.oracle_set_nan_emulation.argtypes = [C.c_int]
        l.oracle_tonemap.restype = C.c_int
        l.oracle_tonemap.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_double,
                                     C.c_void_p, _u64p, _u64p, _f64p, _f64p]
        l.oracle_iterate_points.restype = C.c_int
        l.oracle_iterate_points.argtypes = [C.POINTER(ffr.FfrFlameDesc), C.c_int64, C.c_uint64,
                                            _u64p, _f64p, _f64p]
        _oracle = l
    return _oracle


def ref(elem_size=8):
    """The unmodified reference behind ref_harness.cpp: elem_size 8 = as shipped (double /
    uint64_t), 4 = its float / uint32_t configuration (`make -C oracle ref32`)."""
    if elem_size not in _ref:
        l = C.CDLL(REF_LIB if elem_size == 8 else REF_LIB_F32)
        l.ref_elem_size.restype = C.c_int
        assert l.ref_elem_size() == elem_size
        l.ref_last_error.restype = C.c_char_p
        l.ref_splitmix64.restype = C.c_uint64
        l.ref_splitmix64.argtypes = [C.c_uint64]
        l.ref_isaac_words.argtypes = [C.c_uint64, C.c_uint64, _u64p]
        l.ref_rand_nums.argtypes = [C.c_uint64, C.c_uint64, _f64p]
        l.ref_json_dims.argtypes = [C.c_char_p]
        l.ref_flame_info.argtypes = [C.c_char_p, _u64p, _u64p, _f64p, _f64p, _u64p, _u64p, _u64p]
        l.ref_render_chains.argtypes = [C.c_char_p] + [C.c_uint64] * 6 + [
            C.c_void_p, C.c_uint64, C.POINTER(ffr.FfrStats)]
        l.ref_render_raw.argtypes = [C.c_char_p] + [C.c_uint64] * 4 + [
            C.c_void_p, C.c_uint64, C.POINTER(ffr.FfrStats)]
        l.ref_render_mt.argtypes = [C.c_char_p] + [C.c_uint64] * 4 + [
            C.POINTER(C.c_double), C.c_void_p, C.c_uint64, C.POINTER(ffr.FfrStats)]
        l.ref_iterate_points.argtypes = [C.c_char_p, C.c_int64, C.c_uint64, _u64p, _f64p, _f64p]
        _ref[elem_size] = l
    return _ref[elem_size]


def resize_json(text, size):
    """The reference has no size setter: edit the JSON text.
    Replaces the first uncommented "size" array."""
    if isinstance(text, bytes):
        text = text.decode()
    out, done = [], False
    for line in text.split("\n"):
        if not done and re.match(r'^\s*"size"\s*:', line):
            line = re.sub(r'"size"\s*:\s*\[[^\]]*\]', '"size": %s' % json.dumps(list(size)), line)
            done = True
        out.append(line)
    if not done:
        raise ValueError("no size key found")
    return "\n".join(out)


def _buffer_len(flame):
    _, _, cells, cs = flame.layout()
    return cells * cs


def oracle_render(flame, chain_count, chain_len, base_seed=1, chain_first=0, last_len=0,
                  bv_limit=256, nthreads=1, into=None):
    """Returns (raw uint64 buffer, stats dict, ok)."""
    buf = into if into is not None else np.zeros(_buffer_len(flame), dtype=np.uint64)
    st = ffr.FfrStats()
    rc = oracle().oracle_render_chains(flame.desc_p, base_seed, chain_first, chain_count,
                                       chain_len, last_len, bv_limit,
                                       buf.ctypes.data_as(C.c_void_p), C.byref(st), nthreads)
    if rc < 0:
        raise RuntimeError("oracle_render_chains failed")
    return buf, ffr.stats_to_dict(st, flame.dims, flame.desc.num_xform_ids), rc == 0


def set_nan_emulation(on):
    """See oracle_set_nan_emulation in ffr_oracle.c: only for pinning against oracle/_ref."""
    oracle().oracle_set_nan_emulation(1 if on else 0)


def oracle_render_samples(flame, samples, chain_len, **kw):
    chains = (samples + chain_len - 1) // chain_len
    last = samples - (chains - 1) * chain_len
    return oracle_render(flame, chains, chain_len, last_len=(0 if last == chain_len else last), **kw)


def oracle_iterate_points(flame, xf_index, seeds, pts):
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
    out = np.empty_like(pts)
    rc = oracle().oracle_iterate_points(flame.desc_p, xf_index, len(seeds),
                                        seeds.ctypes.data_as(_u64p), pts.ctypes.data_as(_f64p),
                                        out.ctypes.data_as(_f64p))
    if rc:
        raise RuntimeError("oracle_iterate_points failed")
    return out


def oracle_tonemap(raw, width, height, color_dims, mode, bits=8, gamma=1.0):
    """ffr-img pixel math on the CPU: returns (image, info dict)."""
    raw = np.ascontiguousarray(raw).view(np.uint64)
    ch = 3 if mode == 3 else 1
    if mode == 1:
        bits = 8
    img = np.zeros((height, width, ch), dtype=np.uint8 if bits == 8 else np.uint16)
    hmin, hmax = C.c_uint64(), C.c_uint64()
    smin, smax = C.c_double(), C.c_double()
    rc = oracle().oracle_tonemap(raw.ctypes.data_as(C.c_void_p), width * height, 1 + color_dims,
                                 mode, bits, gamma, img.ctypes.data_as(C.c_void_p),
                                 C.byref(hmin), C.byref(hmax), C.byref(smin), C.byref(smax))
    if rc:
        raise RuntimeError("histogram is (probably) empty")
    info = {"hist_min": hmin.value, "hist_max": hmax.value, "scaler_min": smin.value,
            "scaler_max": smax.value}
    return (img[:, :, 0] if ch == 1 else img), info


def oracle_isaac_words(seed, n):
    out = np.empty(n, dtype=np.uint64)
    oracle().oracle_isaac_words(seed, n, out.ctypes.data_as(_u64p))
    return out


def _ref_err(rc, elem_size=8):
    if rc < 0:
        raise RuntimeError("reference: " + ref(elem_size).ref_last_error().decode())


def _dtype(elem_size):
    return np.uint64 if elem_size == 8 else np.uint32


def _enc(text):
    return text.encode() if isinstance(text, str) else text


def ref_isaac_words(seed, n, elem_size=8):
    out = np.empty(n, dtype=np.uint64)
    ref(elem_size).ref_isaac_words(seed, n, out.ctypes.data_as(_u64p))
    return out


def ref_flame_info(text, elem_size=8):
    n = C.c_uint64()
    ids = (C.c_uint64 * 256)()
    cw = (C.c_double * 256)()
    md = (C.c_double * 3)()
    mi = (C.c_uint64 * 3)()
    cells, cs = C.c_uint64(), C.c_uint64()
    _ref_err(ref(elem_size).ref_flame_info(_enc(text), C.byref(n), ids, cw, md, mi, C.byref(cells),
                                           C.byref(cs)), elem_size)
    dims = ref(elem_size).ref_json_dims(_enc(text))
    k = n.value
    return {"ids": list(ids)[:k], "cw": list(cw)[:k], "mult_d": list(md)[:dims],
            "mult_i": list(mi)[:dims], "cells": cells.value, "cell_size": cs.value, "dims": dims}


def ref_render(text, chain_count, chain_len, base_seed=1, chain_first=0, last_len=0,
               bv_limit=256, elem_size=8):
    info = ref_flame_info(text, elem_size)
    buf = np.zeros(info["cells"] * info["cell_size"], dtype=_dtype(elem_size))
    st = ffr.FfrStats()
    rc = ref(elem_size).ref_render_chains(_enc(text), base_seed, chain_first, chain_count, chain_len,
                                          last_len, bv_limit, buf.ctypes.data_as(C.c_void_p),
                                          buf.nbytes, C.byref(st))
    _ref_err(rc, elem_size)
    n_ids = len(json_xforms_count(text))
    return buf, ffr.stats_to_dict(st, info["dims"], n_ids), rc == 0


def json_xforms_count(text):
    # number of xform ids = entries of the JSON "xforms" array; parsed by our own flame model
    f = ffr.Flame(text)
    return range(f.desc.num_xform_ids)


def ref_render_raw(text, raw_seed, samples, batch, bv_limit=256):
    info = ref_flame_info(text)
    buf = np.zeros(info["cells"] * info["cell_size"], dtype=np.uint64)
    st = ffr.FfrStats()
    rc = ref().ref_render_raw(_enc(text), raw_seed, samples, batch, bv_limit,
                              buf.ctypes.data_as(C.c_void_p), buf.nbytes, C.byref(st))
    _ref_err(rc)
    return buf, ffr.stats_to_dict(st, info["dims"], len(json_xforms_count(text))), rc == 0


def ref_render_mt(text, samples, threads, batch, bv_limit=256, want_buffer=False):
    """BufferRenderer::render() timed: returns (seconds, stats, buffer or None)."""
    info = ref_flame_info(text)
    secs = C.c_double()
    st = ffr.FfrStats()
    buf = np.zeros(info["cells"] * info["cell_size"], dtype=np.uint64) if want_buffer else None
    rc = ref().ref_render_mt(_enc(text), samples, threads, batch, bv_limit, C.byref(secs),
                             buf.ctypes.data_as(C.c_void_p) if want_buffer else None,
                             buf.nbytes if want_buffer else 0, C.byref(st))
    _ref_err(rc)
    return secs.value, ffr.stats_to_dict(st, info["dims"], len(json_xforms_count(text))), buf


def ref_iterate_points(text, xf_index, seeds, pts, elem_size=8):
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
    out = np.empty_like(pts)
    _ref_err(ref(elem_size).ref_iterate_points(_enc(text), xf_index, len(seeds),
                                               seeds.ctypes.data_as(_u64p), pts.ctypes.data_as(_f64p),
                                               out.ctypes.data_as(_f64p)), elem_size)
    return out


_refimg = None


def have_refimg():
    return os.path.exists(REFIMG_LIB)


def refimg():
    global _refimg
    if _refimg is None:
        l = C.CDLL(REFIMG_LIB)
        l.refimg_last_error.restype = C.c_char_p
        l.refimg_render.restype = C.c_int
        l.refimg_render.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_double,
                                    C.c_void_p, C.c_size_t, _u64p, _f64p]
        l.refimg_flame_echo.restype = C.c_size_t
        l.refimg_flame_echo.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
        _refimg = l
    return _refimg


def ref_tonemap(text, raw, mode, bits=8, gamma=1.0):
    """The reference's own render_image() (ffr_img.cpp:199-309 over image_renderer.hpp) on a raw
    double/u64 buffer: returns (image, info) like oracle_tonemap."""
    fl = ffr.Flame(text)
    w, h = fl.size[0], fl.size[1]
    raw = np.ascontiguousarray(raw).view(np.uint64)
    ch = 3 if mode == 3 else 1
    if mode == 1:
        bits = 8
    img = np.zeros((h, w, ch), dtype=np.uint8 if bits == 8 else np.uint16)
    hb = (C.c_uint64 * 2)()
    lb = (C.c_double * 2)()
    rc = refimg().refimg_render(_enc(text), raw.ctypes.data_as(C.c_void_p), raw.nbytes, mode, bits,
                                gamma, img.ctypes.data_as(C.c_void_p), img.nbytes, hb, lb)
    if rc:
        raise RuntimeError("reference: " + refimg().refimg_last_error().decode())
    info = {"hist_min": hb[0], "hist_max": hb[1], "scaler_min_printed": lb[0],
            "scaler_max_printed": lb[1]}
    return (img[:, :, 0] if ch == 1 else img), info


def ref_flame_echo(text):
    n = refimg().refimg_flame_echo(_enc(text), None, 0)
    if not n:
        raise RuntimeError("reference: " + refimg().refimg_last_error().decode())
    out = C.create_string_buffer(n + 1)
    refimg().refimg_flame_echo(_enc(text), out, n + 1)
    return out.value.decode()
