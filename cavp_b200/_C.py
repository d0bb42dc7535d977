"""ctypes binding of libcavp_b200.so - the C ABI declared in include/cavp_b200.h.

Signatures are parsed from the header, so the binding cannot drift from the declaration.  There is no fallback path:
a missing library is built with nvcc (cavp_b200/build.py); if that fails, or a launcher returns non-zero, we raise.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "cavp_b200.h")
LIB_PATH = os.path.join(_HERE, "lib", "libcavp_b200.so")
_lib = None
_fns = {}

_CTYPES = {"int": ctypes.c_int, "float": ctypes.c_float, "double": ctypes.c_double, "long long": ctypes.c_longlong}


class CavpError(RuntimeError):
    pass


def parse_header(path=HEADER):
    """-> {name: [ctypes argtypes]} for every `int cavp_*(...)` prototype."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\bint\s+(cavp_\w+)\s*\(([^)]*)\)\s*;", text):
        name, args = m.group(1), m.group(2)
        argtypes = []
        for a in [x.strip() for x in args.split(",") if x.strip() and x.strip() != "void"]:
            if "*" in a:
                argtypes.append(ctypes.c_void_p)
            else:
                base = re.sub(r"\b(const|unsigned)\b", "", a).strip()
                base = " ".join(base.split()[:-1])  # drop the parameter name
                argtypes.append(_CTYPES[base])
        protos[name] = argtypes
    return protos


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            from . import build as _build
            _build.build()
        _lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in parse_header().items():
            fn = getattr(_lib, name)  # AttributeError here = header / library mismatch
            fn.argtypes = argtypes
            fn.restype = ctypes.c_int
            _fns[name] = fn
    return _lib


def ptr(t):
    return 0 if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    if not _fns:
        lib()
    rc = _fns[name](*args)
    if rc != 0:
        raise CavpError(f"{name} failed with status {rc}")


def query(name, *args):
    """For the few entry points that return a size instead of a status."""
    if not _fns:
        lib()
    return _fns[name](*args)
