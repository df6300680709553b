"""The run-time compiled kernel's build check: source generation + NVRTC need no GPU."""
import pytest

import flames


def test_generate_and_compile_without_device(ffr, examples):
    fl = ffr.Flame(examples.example_json("csci6360_project", size=[256, 256]))
    try:
        src, n = ffr.jit_compile(fl)
    except ffr.FfrError as e:
        if "libnvrtc not found" in str(e):
            pytest.skip("no NVRTC on this machine")
        raise
    assert n > 10000
    for k in range(6):
        assert "jx_%d(" % k in src            # 5 xforms + the final xform, one function each
    assert "__constant__ JT jc[" in src and '#include "ffr_jit_async.cuh"' in src
    # literals are hexadecimal floating point: exactly the blob's values
    assert "0x1.ccccccccccccdp-1" in src      # 0.9


def test_float_3d_with_rng_variations_compiles(ffr):
    text = flames.multi_variation_flame(["julian", "blur", "pie", "disc2"], dims=3, final_name="noise")
    fl = ffr.Flame(text, elem_size=4)
    try:
        src, n = ffr.jit_compile(fl)
    except ffr.FfrError as e:
        if "libnvrtc not found" in str(e):
            pytest.skip("no NVRTC on this machine")
        raise
    assert n > 10000 and "typedef float JT;" in src and "#define JANY_RNG 1" in src


def test_too_many_xforms_is_refused(ffr):
    with pytest.raises(ffr.FfrError):
        ffr.jit_compile(ffr.Flame(flames.many_xforms_flame(20)))


_CACHE_PROBE = r"""
import importlib, os, sys
sys.path.insert(0, sys.argv[1])
ffr = importlib.import_module("flame-fractal-renderer_b200")
ex = importlib.import_module("flame-fractal-renderer_b200.examples")
fl = ffr.Flame(ex.example_json("sierpinski_with_variations", size=[64, 64]))
try:
    ffr.jit_compile(fl)
except ffr.FfrError as e:
    print("SKIP" if "libnvrtc not found" in str(e) else "ERR " + str(e))
    raise SystemExit(0)
print("OK")
"""


def _compile_in_fresh_process(cache_dir):
    """The in-process cache would hide the disk cache: one compile per process."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FFR_JIT_CACHE=str(cache_dir))
    env.pop("FFR_JIT_NO_DISK_CACHE", None)
    p = subprocess.run([sys.executable, "-c", _CACHE_PROBE, root], env=env, capture_output=True,
                       text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    return p.stdout.strip()


def test_disk_cache_only_in_a_private_directory_and_entries_are_verified(tmp_path):
    """a cubin is code that runs in the caller's CUDA context. The cache directory is
    used only when it is a real directory of this user with mode 0700; an entry must carry the
    full key and an intact payload or it is recompiled."""
    import os
    shared = tmp_path / "shared"
    shared.mkdir(mode=0o755)
    os.chmod(shared, 0o755)
    out = _compile_in_fresh_process(shared)
    if out == "SKIP":
        pytest.skip("no NVRTC on this machine")
    assert out == "OK"
    assert list(shared.iterdir()) == []            # group/world-accessible: refused, nothing written
    link_target = tmp_path / "elsewhere"
    link_target.mkdir(mode=0o700)
    link = tmp_path / "link"
    link.symlink_to(link_target)
    assert _compile_in_fresh_process(link) == "OK"
    assert list(link_target.iterdir()) == []       # a symlink is not a directory of ours
    private = tmp_path / "fresh" / "ffr-b200-jit"  # created on first use, 0700
    (tmp_path / "fresh").mkdir()
    assert _compile_in_fresh_process(private) == "OK"
    assert (os.stat(private).st_mode & 0o777) == 0o700
    entries = list(private.iterdir())
    assert len(entries) == 1 and entries[0].name.endswith(".cubin") and len(entries[0].name) == 32 + 6
    assert (os.stat(entries[0]).st_mode & 0o777) == 0o600
    good = entries[0].read_bytes()
    assert good[:7] == b"FFRJIT2"
    # a damaged payload, a truncated file and a foreign file under the right name are all refused
    # and replaced by a fresh compile
    for bad in (good[:200] + bytes([good[200] ^ 1]) + good[201:], good[:-100], b"\x7fELF" + good[4:]):
        entries[0].write_bytes(bad)
        assert _compile_in_fresh_process(private) == "OK"
        assert entries[0].read_bytes() == good
