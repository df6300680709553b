"""f1: the log-density tone map (ffr-img's pixel math, src/ffr_img.cpp:199-309) on the device
against (a) the reference's own code -- render_image() of ffr_img.cpp over
renderers/image_renderer.hpp, compiled by `make -C oracle refimg` into
oracle/_ref/libffr_refimg.so, which travels to the GPU box -- and (b) the oracle restatement,
which tests/test_golden.py and tests/test_oracle_vs_reference.py show to be bit-identical to (a).
Tolerance: log/pow are libm vs CUDA (1-2 ULP), so a pixel may differ by one code when v*scale
falls within ~1e-13 of an integer: at most 1 code, on at most 0.01 % of the pixels; mono mode
and the histogram bounds are exact."""
import io
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rendered(ffr, examples, name, size, samples=3_000_000):
    fl = ffr.Flame(examples.example_json(name, size=size))
    r = ffr.BufferRenderer(fl)
    assert r.render(samples, 1024, base_seed=4)
    return fl, r


def close_enough(a, b):
    d = np.abs(a.astype(np.int64) - b.astype(np.int64))
    assert d.max() <= 1
    assert (d != 0).mean() <= 1e-4


@pytest.mark.parametrize("bits", [8, 16])
@pytest.mark.parametrize("gamma", [1.0, 2.2, 0.5])
def test_gray(ffr, po, examples, bits, gamma):
    fl, r = rendered(ffr, examples, "csci6360_project", [320, 180])
    img, info = r.tonemap(ffr.TONE_GRAY, bits=bits, gamma=gamma)
    raw = r.read_buffer()
    want, oinfo = po.oracle_tonemap(raw, 320, 180, 0, 2, bits=bits, gamma=gamma)
    r.close()
    assert img.shape == (180, 320) and img.dtype == (np.uint8 if bits == 8 else np.uint16)
    assert info["hist_min"] == oinfo["hist_min"] and info["hist_max"] == oinfo["hist_max"]
    assert info["scaler_max"] == oinfo["scaler_max"]  # host glibc log on both sides
    close_enough(img, want)
    # the brightest cell maps to the top code exactly (l == 1.0), empty cells to 0
    counts = raw.reshape(180, 320)
    assert img[counts == counts.max()].min() == (255 if bits == 8 else 65535)
    assert (img[counts == 0] == 0).all()


def test_mono_exact(ffr, po, examples):
    fl, r = rendered(ffr, examples, "sierpinski_triangle", [256, 256], samples=300_000)
    img, _ = r.tonemap(ffr.TONE_MONO)
    raw = r.read_buffer()
    want, _ = po.oracle_tonemap(raw, 256, 256, 0, 1)
    r.close()
    assert np.array_equal(img, want)
    assert set(np.unique(img)) <= {0, 255}


@pytest.mark.parametrize("bits", [8, 16])
def test_rgb(ffr, po, examples, bits):
    fl, r = rendered(ffr, examples, "tkoz_test3", [320, 180])
    img, info = r.tonemap(ffr.TONE_RGB, bits=bits, gamma=2.0)
    raw = r.read_buffer()
    want, _ = po.oracle_tonemap(raw, 320, 180, 3, 3, bits=bits, gamma=2.0)
    r.close()
    assert img.shape == (180, 320, 3)
    close_enough(img, want)


@pytest.mark.parametrize("name,size,cd", [("tkoz_test3", [320, 180], 3),
                                          ("csci6360_project", [256, 144], 0)])
def test_against_reference_compiled_tonemap(ffr, po, examples, name, size, cd):
    """Every mode x bit depth x gamma against the reference's own render_image()."""
    if not po.have_refimg():
        pytest.skip("oracle/_ref/libffr_refimg.so not built")
    text = examples.example_json(name, size=size)
    fl, r = rendered(ffr, examples, name, size)
    raw = r.read_buffer()
    n = 0
    for mode in ((ffr.TONE_MONO, ffr.TONE_GRAY, ffr.TONE_RGB) if cd == 3 else (ffr.TONE_MONO, ffr.TONE_GRAY)):
        for bits in (8, 16):
            for gamma in (1.0, 2.2, 0.5, 4.0):
                img, info = r.tonemap(mode, bits=bits, gamma=gamma)
                want, winfo = po.ref_tonemap(text, raw, mode, bits, gamma)
                assert img.shape == want.shape and img.dtype == want.dtype
                assert info["hist_min"] == winfo["hist_min"] and info["hist_max"] == winfo["hist_max"]
                if mode == ffr.TONE_MONO:
                    assert np.array_equal(img, want)
                else:
                    close_enough(img, want)
                n += 1
    r.close()
    assert n == (24 if cd == 3 else 16)


def test_errors(ffr, examples):
    fl, r = rendered(ffr, examples, "csci6360_project", [64, 36], samples=100_000)
    with pytest.raises(ffr.FfrError, match="buffer must use 3 color dimensions"):
        r.tonemap(ffr.TONE_RGB)
    with pytest.raises(ffr.FfrError, match="gamma too small"):
        r.tonemap(ffr.TONE_GRAY, gamma=0.0)
    with pytest.raises(ffr.FfrError, match="bits per channel must be 8 or 16"):
        r.tonemap(ffr.TONE_GRAY, bits=12)
    r.clear()
    with pytest.raises(ffr.FfrError, match="probably. empty"):
        r.tonemap(ffr.TONE_GRAY)
    r.close()
    fl3 = ffr.Flame(examples.example_json("sierpinski_triangle_3d", size=[16, 16, 16]))
    r3 = ffr.BufferRenderer(fl3)
    with pytest.raises(ffr.FfrError, match="only 2D flames supported"):
        lib_rc = ffr.lib().ffr_cuda_tonemap(r3._h, 2, 8, 1.0, (ffr.C.c_char * 16)(), 16, None)
        r3._check(lib_rc)
    r3.close()


def decode_png(path):
    """Minimal PNG reader for what ffr-img.out writes (filter 0, no interlace)."""
    import struct
    import zlib
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, hdr = 8, b"", None
    while pos < len(data):
        n, typ = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        crc, = struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])
        assert zlib.crc32(typ + body) == crc
        if typ == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif typ == b"IDAT":
            idat += body
        pos += 12 + n
    w, h, bits, ctype = hdr[:4]
    ch = 3 if ctype == 2 else 1
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + w * ch * bits // 8)
    assert (raw[:, 0] == 0).all()
    px = raw[:, 1:]
    if bits == 16:
        px = px.reshape(h, w * ch, 2).astype(np.uint16)
        px = (px[:, :, 0] << 8) | px[:, :, 1]
    px = px.reshape(h, w, ch)
    return px[:, :, 0] if ch == 1 else px


def test_ffr_img_cli_png(ffr, po, examples, tmp_path):
    """ffr-buf.out | ffr-img.out: the buffer file format feeds the image tool unchanged, two -i
    buffers add up, and the PNG decodes to the device pixels."""
    from PIL import Image
    here = os.path.dirname(ffr.LIB_PATH)
    flame = tmp_path / "t3.json"
    flame.write_text(examples.example_json("tkoz_test3", size=[200, 120]))
    buf = tmp_path / "t3.buf"
    p = subprocess.run([os.path.join(here, "ffr-buf.out"), "-f", str(flame), "-o", str(buf),
                        "-s", "2000000", "-b", "1000", "--seed", "9"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    raw = np.fromfile(buf, dtype=np.uint64)
    for flags, mode, bits, cd in ((["-c", "-b", "16", "-y", "2.0"], 3, 16, 3), (["-g"], 2, 8, 3)):
        png = tmp_path / "out.png"
        p = subprocess.run([os.path.join(here, "ffr-img.out"), "-f", str(flame), "-i", str(buf),
                            "-i", str(buf), "-o", str(png)] + flags, capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        assert "histogram bounds:" in p.stderr and "scaler bounds:" in p.stderr
        got = decode_png(png)
        if bits == 8:
            assert np.array_equal(np.array(Image.open(png)), got)  # a standard decoder agrees
        # two identical inputs: counts and colour sums double
        summed = raw.reshape(-1, 4).copy()
        summed[:, 0] *= 2
        summed[:, 1:] = (summed[:, 1:].view(np.float64) * 2).view(np.uint64)
        gamma = 2.0 if mode == 3 else 1.0
        want, _ = po.oracle_tonemap(summed.ravel(), 200, 120, cd, mode, bits=bits, gamma=gamma)
        assert got.shape == want.shape
        close_enough(got, want)
