"""flame-fractal-renderer_b200: Python host binding over the C ABI of libffr_cuda.

The product is the CUDA library (csrc/, include/ffr_cuda.h) plus the C++ host
(host/: flame model and the ffr-buf.out command line). This module is the thin
ctypes mirror of that ABI used by tests/, bench.py and __graft_entry__.py; it mirrors
the reference's BufferRenderer surface (src/renderers/buffer_renderer.hpp:254-584 in
the reference repo) method for method:

    Flame(json_text)                      Flame<dims>(json)            types/flame.hpp:91
    BufferRenderer(flame)                 BufferRenderer(flame)        buffer_renderer.hpp:254
      .add_buffer(bytes/ndarray)          addBuffer                    :375-452
      .render(samples, chain_len, seed)   render / renderSeeded        :269-372
      .read_buffer() -> ndarray           writeBuffer                  :476-480
      .stats                              getSamplesIterated & co.     :534-564

There is no CPU fallback: importing works anywhere (the library links cudart
statically), but constructing a BufferRenderer without an sm_100 GPU raises.

The directory name contains '-', so import it with
    importlib.import_module("flame-fractal-renderer_b200")
(tests/conftest.py and bench.py do exactly that).
"""

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libffr_cuda.so")

FFR_MAX_DIMS = 3
FFR_MAX_COLOR_DIMS = 127
FFR_MAX_XFORMS = 64
FFR_MAX_VAR_PARAMS = 8
FFR_MAX_BAD_RECORDED = 1024
FFR_FINAL_XFORM_ID = (1 << 64) - 1

FFR_OK = 0
FFR_BAD_VALUES = 1
FFR_E_INVALID = -1
FFR_E_CUDA = -2
FFR_E_NODEVICE = -3
FFR_E_UNSUPPORTED = -4

SCATTER_AUTO, SCATTER_GLOBAL, SCATTER_WARP_AGG, SCATTER_SMEM_TILE, SCATTER_DISCARD = 0, 1, 2, 3, 4


class FfrVariation(C.Structure):
    _fields_ = [("op", C.c_uint32), ("axis_x", C.c_uint32), ("axis_y", C.c_uint32),
                ("reserved", C.c_uint32), ("weight", C.c_double),
                ("params", C.c_double * FFR_MAX_VAR_PARAMS)]


class FfrXForm(C.Structure):
    _fields_ = [("id", C.c_uint64), ("weight", C.c_double),
                ("has_pre", C.c_uint32), ("has_post", C.c_uint32),
                ("has_color", C.c_uint32), ("num_vars", C.c_uint32),
                ("pre_A", C.c_double * 9), ("pre_b", C.c_double * 3),
                ("post_A", C.c_double * 9), ("post_b", C.c_double * 3),
                ("color_speed", C.c_double),
                ("color", C.POINTER(C.c_double)),
                ("vars", C.POINTER(FfrVariation))]


class FfrFlameDesc(C.Structure):
    _fields_ = [("dims", C.c_uint32), ("color_dims", C.c_uint32),
                ("elem_size", C.c_uint32), ("has_final", C.c_uint32),
                ("size", C.c_uint64 * FFR_MAX_DIMS),
                ("bounds_lo", C.c_double * FFR_MAX_DIMS),
                ("bounds_hi", C.c_double * FFR_MAX_DIMS),
                ("num_xforms", C.c_uint32), ("num_xform_ids", C.c_uint32),
                ("xforms", C.POINTER(FfrXForm)),
                ("xfcw", C.POINTER(C.c_double)),
                ("final_xform", C.POINTER(FfrXForm))]


class FfrStats(C.Structure):
    _fields_ = [("s_iter", C.c_uint64), ("s_plot", C.c_uint64),
                ("xf_dist", C.c_uint64 * FFR_MAX_XFORMS),
                ("pt_min", C.c_double * FFR_MAX_DIMS),
                ("pt_max", C.c_double * FFR_MAX_DIMS),
                ("n_bad", C.c_uint64),
                ("bad_xf", C.c_uint64 * FFR_MAX_BAD_RECORDED),
                ("bad_pt", (C.c_double * FFR_MAX_DIMS) * FFR_MAX_BAD_RECORDED)]


class FfrOptions(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("scatter_mode", C.c_uint32),
                ("regroup", C.c_uint32), ("blocks_per_sm", C.c_uint32),
                ("external_buffer", C.c_void_p), ("stream", C.c_void_p),
                ("jit", C.c_uint32), ("reserved0", C.c_uint32)]


JIT_AUTO, JIT_OFF, JIT_ON = 0, 1, 2


class FfrJitInfo(C.Structure):
    _fields_ = [("active", C.c_uint32), ("eligible", C.c_uint32), ("failed", C.c_uint32),
                ("from_cache", C.c_uint32), ("threads_per_block", C.c_uint32),
                ("slots_per_block", C.c_uint32), ("blocks_per_sm", C.c_uint32),
                ("registers", C.c_uint32), ("smem_bytes", C.c_uint64),
                ("cubin_bytes", C.c_uint64), ("source_bytes", C.c_uint64),
                ("compile_seconds", C.c_double), ("message", C.c_char * 256)]


class FfrTonemapInfo(C.Structure):
    _fields_ = [("hist_min", C.c_uint64), ("hist_max", C.c_uint64),
                ("scaler_min", C.c_double), ("scaler_max", C.c_double),
                ("width", C.c_uint32), ("height", C.c_uint32),
                ("channels", C.c_uint32), ("bits", C.c_uint32)]


TONE_MONO, TONE_GRAY, TONE_RGB = 1, 2, 3

PROGRESS_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_uint64, C.c_uint64)

# every symbol include/ffr_cuda.h and include/ffr_flame.h declare: (restype, argtypes)
_u64p = C.POINTER(C.c_uint64)
_f64p = C.POINTER(C.c_double)
_descp = C.POINTER(FfrFlameDesc)
ABI = {
    # ffr_flame.h
    "ffr_flame_from_json": (C.c_void_p, [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]),
    "ffr_flame_from_json_sized": (C.c_void_p, [C.c_char_p, C.c_size_t, _u64p, C.c_int,
                                               C.c_char_p, C.c_size_t]),
    "ffr_flame_from_json_ex": (C.c_void_p, [C.c_char_p, C.c_size_t, _u64p, C.c_int, C.c_int,
                                            C.c_char_p, C.c_size_t]),
    "ffr_flame_get_desc": (_descp, [C.c_void_p]),
    "ffr_flame_free": (None, [C.c_void_p]),
    "ffr_flame_layout": (C.c_int, [_descp, _f64p, _u64p, _u64p, _u64p]),
    "ffr_var_name": (C.c_char_p, [C.c_uint32]),
    "ffr_var_op_from_name": (C.c_uint32, [C.c_char_p]),
    "ffr_reference_batch_size": (C.c_uint64, [C.c_uint64]),
    "ffr_flame_json_echo": (C.c_size_t, [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t,
                                         C.c_char_p, C.c_size_t]),
    # ffr_cuda.h
    "ffr_cuda_version": (C.c_char_p, []),
    "ffr_cuda_device_count": (C.c_int, []),
    "ffr_chain_seed": (C.c_uint64, [C.c_uint64, C.c_uint64]),
    "ffr_cuda_create": (C.c_void_p, [_descp, C.POINTER(C.c_int), C.c_int, C.c_char_p, C.c_size_t]),
    "ffr_cuda_create_ex": (C.c_void_p, [_descp, C.POINTER(C.c_int), C.c_int,
                                        C.POINTER(FfrOptions), C.c_char_p, C.c_size_t]),
    "ffr_cuda_destroy": (None, [C.c_void_p]),
    "ffr_cuda_last_error": (C.c_char_p, [C.c_void_p]),
    "ffr_cuda_buffer_bytes": (C.c_size_t, [C.c_void_p]),
    "ffr_cuda_buffer_cells": (C.c_uint64, [C.c_void_p]),
    "ffr_cuda_device_buffer": (C.c_void_p, [C.c_void_p, C.c_int]),
    "ffr_cuda_add_buffer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ffr_cuda_clear_buffer": (C.c_int, [C.c_void_p]),
    "ffr_cuda_render": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                  PROGRESS_CB, C.c_void_p, C.POINTER(FfrStats)]),
    "ffr_cuda_render_chains": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64,
                                         C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(FfrStats)]),
    "ffr_cuda_render_chains_async": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64,
                                               C.c_uint64, C.c_uint64, C.c_uint64]),
    "ffr_cuda_clear_buffer_async": (C.c_int, [C.c_void_p]),
    "ffr_cuda_add_buffer_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ffr_cuda_read_buffer_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ffr_cuda_sync": (C.c_int, [C.c_void_p]),
    "ffr_cuda_get_stats": (C.c_int, [C.c_void_p, C.POINTER(FfrStats)]),
    "ffr_cuda_resident_chains": (C.c_uint64, [C.c_void_p]),
    "ffr_cuda_launch_count": (C.c_uint64, [C.c_void_p]),
    "ffr_cuda_reduce": (C.c_int, [C.c_void_p]),
    "ffr_cuda_sum_device_slices": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int,
                                             C.c_uint64, C.c_uint64]),
    "ffr_cuda_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ffr_cuda_ipc_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "ffr_cuda_read_buffer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ffr_cuda_histogram_sum_max": (C.c_int, [C.c_void_p, _u64p, _u64p]),
    "ffr_cuda_tonemap": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_size_t,
                                   C.POINTER(FfrTonemapInfo)]),
    "ffr_cuda_iterate_points": (C.c_int, [C.c_void_p, C.c_int64, C.c_uint64, _u64p, _f64p, _f64p]),
    "ffr_cuda_isaac_words": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, _u64p]),
    "ffr_cuda_jit_info": (C.c_int, [C.c_void_p, C.POINTER(FfrJitInfo)]),
    "ffr_cuda_jit_enable": (C.c_int, [C.c_void_p]),
    "ffr_cuda_jit_source": (C.c_size_t, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "ffr_cuda_jit_compile": (C.c_int, [_descp, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t),
                                       C.c_char_p, C.c_size_t]),
    "ffr_cuda_atomic_roofline": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_float)]),
    "ffr_cuda_atomic_roofline_ex": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_float),
                                              _u64p]),
}

_lib = None


def lib():
    """Load libffr_cuda.so (in-tree build). Fails loudly if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libffr_cuda.so is not built (%s): run `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C flame-fractal-renderer_b200`; there is no CPU fallback"
                % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in ABI.items():
            fn = getattr(l, name)  # AttributeError if the ABI symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


class FfrError(RuntimeError):
    pass


class Flame:
    """Flame<dims>(json): parse + validate + flatten a flame JSON text (FLAME_JSON.md)."""

    def __init__(self, text, size=None, elem_size=8):
        """elem_size 8: the shipped double/uint64_t build; 4: the float/uint32_t build."""
        if isinstance(text, str):
            text = text.encode()
        err = C.create_string_buffer(512)
        arr = (C.c_uint64 * len(size))(*size) if size is not None else None
        h = lib().ffr_flame_from_json_ex(text, len(text), arr, len(size) if size is not None else 0,
                                         elem_size, err, len(err))
        if not h:
            raise FfrError(err.value.decode())
        self._h = h
        self.desc_p = lib().ffr_flame_get_desc(h)
        self.desc = self.desc_p.contents

    @classmethod
    def from_file(cls, path, size=None, elem_size=8):
        with open(path, "rb") as f:
            return cls(f.read(), size=size, elem_size=elem_size)

    @property
    def elem_size(self):
        return self.desc.elem_size

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.ffr_flame_free(h)

    @property
    def dims(self):
        return self.desc.dims

    @property
    def color_dims(self):
        return self.desc.color_dims

    @property
    def size(self):
        return [self.desc.size[i] for i in range(self.dims)]

    @property
    def xform_ids(self):
        return [self.desc.xforms[i].id for i in range(self.desc.num_xforms)]

    @property
    def cumulative_weights(self):
        return [self.desc.xfcw[i] for i in range(self.desc.num_xforms)]

    def layout(self):
        """(mult_d, mult_i, cells, cell_size) as BufferRenderer::_init computes them."""
        md = (C.c_double * 3)()
        mi = (C.c_uint64 * 3)()
        cells = C.c_uint64()
        cs = C.c_uint64()
        rc = lib().ffr_flame_layout(self.desc_p, md, mi, C.byref(cells), C.byref(cs))
        if rc != FFR_OK:
            raise FfrError("BufferRenderer(): histogram too big")
        d = self.dims
        return list(md)[:d], list(mi)[:d], cells.value, cs.value

    def uses_only(self, ops):
        ops = set(ops)
        xfs = [self.desc.xforms[i] for i in range(self.desc.num_xforms)]
        if self.desc.has_final:
            xfs.append(self.desc.final_xform.contents)
        return all(xf.vars[k].op in ops for xf in xfs for k in range(xf.num_vars))


def flame_json_echo(text):
    """The text after `flame: ` on the reference's stderr (ffr_buf.cpp:129)."""
    if isinstance(text, str):
        text = text.encode()
    err = C.create_string_buffer(512)
    n = lib().ffr_flame_json_echo(text, len(text), None, 0, err, len(err))
    if not n:
        raise FfrError(err.value.decode())
    out = C.create_string_buffer(n + 1)
    lib().ffr_flame_json_echo(text, len(text), out, n + 1, None, 0)
    return out.value.decode()


def stats_to_dict(st, dims, n_ids):
    nb = min(st.n_bad, FFR_MAX_BAD_RECORDED)
    return {
        "s_iter": st.s_iter, "s_plot": st.s_plot,
        "xf_dist": [st.xf_dist[i] for i in range(n_ids)],
        "pt_min": [st.pt_min[i] for i in range(dims)],
        "pt_max": [st.pt_max[i] for i in range(dims)],
        "n_bad": st.n_bad,
        "bad_xf": [st.bad_xf[i] for i in range(nb)],
        "bad_pt": [[st.bad_pt[i][d] for d in range(dims)] for i in range(nb)],
    }


class BufferRenderer:
    """BufferRenderer<dims> on the GPU(s) through the C ABI."""

    def __init__(self, flame, devices=None, scatter_mode=SCATTER_AUTO, regroup=0,
                 blocks_per_sm=0, external_buffer=None, stream=None, jit=JIT_AUTO):
        self.flame = flame
        L = lib()
        if devices is None:
            devices = [0]
        devs = (C.c_int * len(devices))(*devices)
        opt = FfrOptions(C.sizeof(FfrOptions), scatter_mode, regroup, blocks_per_sm,
                         external_buffer, stream, jit, 0)
        err = C.create_string_buffer(512)
        self._h = L.ffr_cuda_create_ex(flame.desc_p, devs, len(devices), C.byref(opt),
                                       err, len(err))
        if not self._h:
            raise FfrError(err.value.decode())
        self.bytes = L.ffr_cuda_buffer_bytes(self._h)
        self.cells = L.ffr_cuda_buffer_cells(self._h)
        self.elem_size = flame.elem_size
        self.word_dtype = np.uint64 if self.elem_size == 8 else np.uint32
        self.cell_size = 1 + flame.color_dims
        self.own_stream = stream is None
        self._stats = FfrStats()

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.ffr_cuda_destroy(h)

    __del__ = close

    def _check(self, rc):
        if rc < 0:
            raise FfrError("libffr_cuda error %d: %s"
                           % (rc, lib().ffr_cuda_last_error(self._h).decode()))
        return rc

    def add_buffer(self, buf):
        a = np.ascontiguousarray(buf).view(np.uint8).ravel()
        self._check(lib().ffr_cuda_add_buffer(self._h, a.ctypes.data_as(C.c_void_p), a.size))

    def clear(self):
        self._check(lib().ffr_cuda_clear_buffer(self._h))

    def render(self, samples, chain_len, base_seed=1, bv_limit=256, progress=None):
        """render(): True on success, False when the bad value limit was exceeded."""
        cb = PROGRESS_CB(lambda u, d, t: progress(d, t)) if progress else PROGRESS_CB()
        rc = self._check(lib().ffr_cuda_render(self._h, samples, chain_len, base_seed, bv_limit,
                                               cb, None, C.byref(self._stats)))
        return rc == FFR_OK

    def render_chains(self, chain_first, chain_count, chain_len, last_len=0, base_seed=1,
                      bv_limit=256):
        rc = self._check(lib().ffr_cuda_render_chains(
            self._h, chain_first, chain_count, chain_len, last_len, base_seed, bv_limit,
            C.byref(self._stats)))
        return rc == FFR_OK

    def render_chains_async(self, chain_first, chain_count, chain_len, last_len=0, base_seed=1,
                            bv_limit=256):
        self._check(lib().ffr_cuda_render_chains_async(
            self._h, chain_first, chain_count, chain_len, last_len, base_seed, bv_limit))

    # the streaming interface: enqueue only; page-locked host buffers, valid until sync()
    def clear_async(self):
        self._check(lib().ffr_cuda_clear_buffer_async(self._h))

    def add_buffer_async(self, pinned):
        a = np.asarray(pinned)
        assert a.flags.c_contiguous
        self._check(lib().ffr_cuda_add_buffer_async(self._h, a.ctypes.data_as(C.c_void_p), a.nbytes))

    def read_buffer_async(self, pinned_out):
        assert pinned_out.flags.c_contiguous
        self._check(lib().ffr_cuda_read_buffer_async(self._h, pinned_out.ctypes.data_as(C.c_void_p),
                                                     pinned_out.nbytes))

    def sync(self):
        self._check(lib().ffr_cuda_sync(self._h))

    def fetch_stats(self):
        self._check(lib().ffr_cuda_get_stats(self._h, C.byref(self._stats)))
        return self.stats

    @property
    def stats(self):
        return stats_to_dict(self._stats, self.flame.dims, self.flame.desc.num_xform_ids)

    @property
    def resident_chains(self):
        return lib().ffr_cuda_resident_chains(self._h)

    @property
    def launches(self):
        return lib().ffr_cuda_launch_count(self._h)

    def reduce(self):
        self._check(lib().ffr_cuda_reduce(self._h))

    def ipc_export(self):
        """64-byte CUDA IPC handle of this context's buffer (for another process on the box)."""
        h = C.create_string_buffer(64)
        self._check(lib().ffr_cuda_ipc_export(self._h, h))
        return h.raw

    def ipc_add(self, handles):
        """Add the buffers behind other processes' IPC handles into this context's buffer."""
        blob = b"".join(handles)
        self._check(lib().ffr_cuda_ipc_add(self._h, blob, len(handles)))

    def sum_device_slices(self, dst_ptr, src_ptrs, first_elem, n_elems):
        """dst += sum(srcs) on the device, typed by position in the cell (K2d)."""
        arr = (C.c_void_p * len(src_ptrs))(*src_ptrs)
        self._check(lib().ffr_cuda_sum_device_slices(self._h, dst_ptr, arr, len(src_ptrs),
                                                     first_elem, n_elems))

    @property
    def jit_info(self):
        """State of the run-time compiled flame-specialised kernel (ffr_jit_info)."""
        info = FfrJitInfo()
        self._check(lib().ffr_cuda_jit_info(self._h, C.byref(info)))
        d = {k: getattr(info, k) for k, _ in FfrJitInfo._fields_}
        d["message"] = d["message"].decode(errors="replace")
        return d

    def jit_enable(self):
        self._check(lib().ffr_cuda_jit_enable(self._h))

    @property
    def jit_source(self):
        n = lib().ffr_cuda_jit_source(self._h, None, 0)
        buf = C.create_string_buffer(n + 1)
        lib().ffr_cuda_jit_source(self._h, buf, n + 1)
        return buf.value.decode()

    def device_buffer(self, dev_index=0):
        return lib().ffr_cuda_device_buffer(self._h, dev_index)

    def read_buffer(self, out=None):
        """writeBuffer(): the raw buffer, cells x (1 + color_dims) elements, as uint64 (double
        build) or uint32 (float build) words."""
        if out is None:
            out = np.empty(self.bytes // self.elem_size, dtype=self.word_dtype)
        self._check(lib().ffr_cuda_read_buffer(self._h, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def histogram_sum_max(self):
        s, m = C.c_uint64(), C.c_uint64()
        self._check(lib().ffr_cuda_histogram_sum_max(self._h, C.byref(s), C.byref(m)))
        return s.value, m.value

    def tonemap(self, mode, bits=8, gamma=1.0):
        """ffr-img's pixel math on the device: returns (image ndarray [h, w(, 3)], info dict)."""
        w, h = self.flame.size[0], self.flame.size[1]
        ch = 3 if mode == TONE_RGB else 1
        if mode == TONE_MONO:
            bits = 8
        img = np.empty((h, w, ch), dtype=np.uint8 if bits == 8 else np.uint16)
        info = FfrTonemapInfo()
        self._check(lib().ffr_cuda_tonemap(self._h, mode, bits, gamma, img.ctypes.data_as(C.c_void_p),
                                           img.nbytes, C.byref(info)))
        d = {k: getattr(info, k) for k, _ in FfrTonemapInfo._fields_}
        return (img[:, :, 0] if ch == 1 else img), d

    def iterate_points(self, xf_index, seeds, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        out = np.empty_like(pts)
        self._check(lib().ffr_cuda_iterate_points(
            self._h, xf_index, len(seeds), seeds.ctypes.data_as(_u64p),
            pts.ctypes.data_as(_f64p), out.ctypes.data_as(_f64p)))
        return out

    def isaac_words(self, seed, n):
        out = np.empty(n, dtype=np.uint64)
        self._check(lib().ffr_cuda_isaac_words(self._h, seed, n, out.ctypes.data_as(_u64p)))
        return out

    def atomic_roofline(self, n_atomics, pattern=0):
        """(milliseconds, cells hit): pattern 0 uniform random cells, 1 attractor replay."""
        ms = C.c_float()
        n = C.c_uint64()
        self._check(lib().ffr_cuda_atomic_roofline_ex(self._h, n_atomics, pattern, C.byref(ms),
                                                      C.byref(n)))
        return ms.value, n.value


def jit_compile(flame):
    """Generate and compile the flame-specialised kernel without a device: (source, cubin
    bytes). Raises FfrError with the NVRTC log on failure."""
    src = C.create_string_buffer(1 << 20)
    err = C.create_string_buffer(8192)
    n = C.c_size_t()
    rc = lib().ffr_cuda_jit_compile(flame.desc_p, src, len(src), C.byref(n), err, len(err))
    if rc != FFR_OK:
        raise FfrError("jit_compile: %s" % err.value.decode(errors="replace"))
    return src.value.decode(), n.value


def split_counts_colors(raw, cells, color_dims):
    """Split a raw reference-layout buffer into (counts, colour sums): u64/f64 for the double
    build, u32/f32 for the float build (decided by the array's word size)."""
    raw = np.asarray(raw)
    wd, fd = (np.uint64, np.float64) if raw.dtype.itemsize == 8 else (np.uint32, np.float32)
    a = raw.view(wd).reshape(cells, 1 + color_dims)
    counts = a[:, 0].copy()
    colors = a[:, 1:].copy().view(fd) if color_dims else None
    return counts, colors
