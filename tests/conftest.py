import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


# Tests that leave ffr_options.jit on auto mean to exercise the ahead-of-time kernels: do not let
# auto mode pick up a cubin that an earlier JIT test of the same flame left in the cache
# (tests/test_gpu_jit.py::test_auto_mode_is_lazy turns this back on for itself).
os.environ.setdefault("FFR_JIT_USE_CACHED", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ffr():
    return importlib.import_module("flame-fractal-renderer_b200")


@pytest.fixture(scope="session")
def examples():
    return importlib.import_module("flame-fractal-renderer_b200.examples")


@pytest.fixture(scope="session")
def po():
    import pyoracle
    return pyoracle
