"""CPU-side checks of the C ABI: the library builds, loads, and exports every function include/b200_fwdsim.h
declares; with no GPU present compute entry points fail loudly (no fallback)."""
import ctypes as C

import numpy as np
import pytest

from pygsti_b200 import _lib, build


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_every_declared_function_is_exported_and_bound(lib):
    declared = _lib.header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "header declares %s but the library does not export it" % name
        assert name in _lib.SIGNATURES, "no ctypes signature for %s" % name
    assert sorted(_lib.SIGNATURES) == declared


def test_version_and_error_string(lib):
    assert lib.b200_version() >= 100
    assert isinstance(lib.b200_last_error(), bytes)


def test_no_silent_cpu_fallback(lib):
    n = C.c_int(-1)
    rc = lib.b200_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p(0)
    rc = lib.b200_ctx_create(0, None, C.byref(h))
    assert rc == _lib.E_CUDA and not h.value
    assert b"no CPU fallback" in lib.b200_last_error()
    with pytest.raises(_lib.B200Error):
        from pygsti_b200 import engine
        engine.Context(0)
