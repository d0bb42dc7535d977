"""ctypes binding of libcavp_b200.so (the C-ABI declared in include/cavp_b200.h).

The product path has no fallback: if the library is missing it is built with nvcc; if that fails, or a launcher
returns non-zero, we raise.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "lib", "libcavp_b200.so")
_lib = None

c_int, c_float, c_void_p, c_ll = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_longlong
P, I, F, L = c_void_p, c_int, c_float, c_ll

# name -> argtypes (return type is always int status)
SIGNATURES = {
    "cavp_igemm": [P] * 8 + [I] * 19 + [I, F, I, I, P],
    "cavp_igemm_wgrad": [P] * 3 + [I] * 16 + [P],
}


class CavpError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            from . import build as _build
            _build.build()
        _lib = ctypes.CDLL(_LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.argtypes = argtypes
            fn.restype = c_int
    return _lib


def ptr(t):
    return 0 if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise CavpError(f"{name} failed with status {rc}")
