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
