"""The C-ABI library loads, exports every symbol include/msda_b200.h declares, and rejects bad
arguments with the documented codes -- all without touching a GPU."""
import ctypes
import os
import re

import pytest

from detrpose_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "msda_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    return re.findall(r"MSDA_API\s+[\w\s\*]+?\b(msda_b200_\w+)\s*\(", text)


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.load()


def test_header_declares_the_path():
    names = declared_symbols()
    for required in ("msda_b200_forward", "msda_b200_backward", "msda_b200_sample_indices",
                     "msda_b200_locations", "msda_b200_repack", "msda_b200_unpack_grad",
                     "msda_b200_last_error", "msda_b200_abi_version"):
        assert required in names


def test_every_declared_symbol_is_exported(lib):
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
    # and the Python binding covers exactly the declared set
    assert sorted(_lib.SIGNATURES) == sorted(set(declared_symbols()))


def test_abi_version(lib):
    assert lib.msda_b200_abi_version() == 1


def test_header_constants_match_binding():
    text = open(HEADER).read()
    consts = dict(re.findall(r"#define\s+(MSDA_\w+)\s+(-?\d+)", text))
    assert int(consts["MSDA_F32"]) == _lib.F32 and int(consts["MSDA_BF16"]) == _lib.BF16
    assert int(consts["MSDA_COORD_UNFUSED"]) == _lib.COORD_UNFUSED
    assert int(consts["MSDA_COORD_FMA"]) == _lib.COORD_FMA
    assert int(consts["MSDA_MAX_LEVELS"]) == _lib.MAX_LEVELS
    assert int(consts["MSDA_MAX_POINTS"]) == _lib.MAX_POINTS


def test_argument_errors_without_gpu(lib):
    shapes = _lib.i32_array([4, 4])
    strides = _lib.i64_array([16 * 8, 8, 8])
    # NULL spatial shapes
    rc = lib.msda_b200_forward(None, 0, strides, None, None, None, None, 0, 1, 1, 1, 8, 1, 1, 0, None)
    assert rc == -1 and b"spatial_shapes" in lib.msda_b200_last_error()
    # non-positive size
    rc = lib.msda_b200_forward(None, 0, strides, shapes, None, None, None, 0, 0, 1, 1, 8, 1, 1, 0, None)
    assert rc == -2
    # unsupported head dim
    rc = lib.msda_b200_forward(None, 0, strides, shapes, None, None, None, 0, 1, 1, 1, 12, 1, 1, 0, None)
    assert rc == -4 and b"Dh=12" in lib.msda_b200_last_error()
    # too many levels / points
    rc = lib.msda_b200_forward(None, 0, strides, shapes, None, None, None, 0, 1, 1, 1, 8, 9, 1, 0, None)
    assert rc == -6
    # NULL value
    rc = lib.msda_b200_forward(None, 0, strides, shapes, None, None, None, 0, 1, 1, 1, 8, 1, 1, 0, None)
    assert rc == -1
    # misaligned value pointer
    rc = lib.msda_b200_forward(ctypes.c_void_p(0x1004), 0, strides, shapes, None, None, None, 0,
                               1, 1, 1, 8, 1, 1, 0, None)
    assert rc == -5
    # unknown dtype
    rc = lib.msda_b200_forward(ctypes.c_void_p(0x1000), 7, strides, shapes, None, None, None, 0,
                               1, 1, 1, 8, 1, 1, 0, None)
    assert rc == -3
    # backward: grads must come together
    rc = lib.msda_b200_backward(ctypes.c_void_p(0x1000), 0, strides, shapes, ctypes.c_void_p(0x1000),
                                ctypes.c_void_p(0x1000), ctypes.c_void_p(0x1000), 0, None, 0,
                                ctypes.c_void_p(0x1000), None, 1, 1, 1, 8, 1, 1, 0, None)
    assert rc == -1
    # prologue: ref_levels must be 1 or L
    rc = lib.msda_b200_locations(ctypes.c_void_p(0x1000), ctypes.c_void_p(0x1000), ctypes.c_void_p(0x1000), 3,
                                 shapes, ctypes.c_void_p(0x1000), ctypes.c_void_p(0x1000), 1, 1, 1, 1, 1, None)
    assert rc == -2


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.MSDAError, match="no CPU / PyTorch fallback"):
        _lib.load()
