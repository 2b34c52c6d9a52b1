"""ctypes binding of ``libmsda_b200.so`` (C ABI in ``include/msda_b200.h``).

The shared library is built in-tree by ``__graft_entry__.build()`` (or
``make -C detrpose_b200/csrc``).  There is no CPU or PyTorch fallback: if the
library is missing or a call fails this module raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmsda_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

F32, BF16 = 0, 1
COORD_UNFUSED, COORD_FMA = 0, 1
MAX_LEVELS, MAX_POINTS = 8, 16

# every symbol include/msda_b200.h declares: (restype, argtypes)
_I32P = POINTER(c_int32)
_I64P = POINTER(c_int64)
_VPP = POINTER(c_void_p)
SIGNATURES = {
    "msda_b200_abi_version": (c_int, []),
    "msda_b200_last_error": (c_char_p, []),
    "msda_b200_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "msda_b200_set_variant": (c_int, [c_int, c_int]),
    "msda_b200_debug_phase_buffer": (c_int, [c_void_p]),
    "msda_b200_forward": (c_int, [c_void_p, c_int, _I64P, _I32P, c_void_p, c_void_p, c_void_p, c_int,
                                  c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "msda_b200_backward": (c_int, [c_void_p, c_int, _I64P, _I32P, c_void_p, c_void_p, c_void_p, c_int,
                                   c_void_p, c_int, c_void_p, c_void_p,
                                   c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "msda_b200_fused_supported": (c_int, [c_int, _I64P, _I32P, c_int, c_int, c_int, c_int, c_int, c_int]),
    "msda_b200_forward_fused": (c_int, [c_void_p, c_int, _I64P, _I32P, c_void_p, c_void_p, c_void_p, c_int,
                                        c_void_p, c_int, c_void_p,
                                        c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "msda_b200_backward_fused": (c_int, [c_void_p, c_int, _I64P, _I32P, c_void_p, c_void_p, c_int, c_void_p,
                                         c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                         c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "msda_b200_softmax_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "msda_b200_sample_indices": (c_int, [_I32P, c_void_p, c_void_p, c_void_p,
                                         c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "msda_b200_locations": (c_int, [c_void_p, c_void_p, c_void_p, c_int, _I32P, c_void_p, c_void_p,
                                    c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "msda_b200_repack": (c_int, [_VPP, _I64P, c_int, _I32P, c_void_p, c_int,
                                 c_int, c_int, c_int, c_int, c_void_p]),
    "msda_b200_unpack_grad": (c_int, [c_void_p, _I32P, _VPP, _I64P, c_int,
                                      c_int, c_int, c_int, c_int, c_void_p]),
    "msda_b200_gate_forward": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_float,
                                       c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "msda_b200_gate_backward": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                        c_void_p]),
    "msda_b200_lqe_forward": (c_int, [c_void_p, c_int, _I64P, c_void_p, c_void_p, c_void_p,
                                      c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "msda_b200_lqe_backward": (c_int, [c_void_p, c_int, _I64P, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
}

_lock = threading.Lock()
_lib = None


class MSDAError(RuntimeError):
    """A C-ABI call returned non-zero."""


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into ``libmsda_b200.so`` (in-tree)."""
    proc = subprocess.run(["make", "-C", CSRC_DIR, "-j", str(os.cpu_count() or 4)],
                          capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        print(proc.stdout)
        print(proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("building libmsda_b200.so failed (see output above)")
    return LIB_PATH


def load():
    """Load the library once and attach the prototypes.  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise MSDAError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C detrpose_b200/csrc` (there is no CPU / PyTorch fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the symbol is missing
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.msda_b200_abi_version() != 1:
            raise MSDAError("libmsda_b200.so ABI version mismatch")
        _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().msda_b200_last_error().decode("utf-8", "replace")
        raise MSDAError(f"{what} failed with code {rc}: {msg}")


def i32_array(values):
    values = [int(v) for v in values]
    return (c_int32 * len(values))(*values)


def i64_array(values):
    values = [int(v) for v in values]
    return (c_int64 * len(values))(*values)


def ptr_array(values):
    return (c_void_p * len(values))(*[int(v) for v in values])
