"""ctypes binding of the C-ABI kernel library (include/camradepth_b200.h).

The prototypes are parsed from the header itself, so the binding cannot drift from the ABI.
There is NO fallback: if the shared library is missing, or a CUDA call is attempted without it,
an error is raised (the product path has no CPU route of any kind).
"""
from __future__ import annotations

import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "camradepth_b200.h")
LIB_PATH = os.environ.get("CAMRADEPTH_LIB") or os.path.join(HERE, "libcamradepth_b200.so")   # override: A/B of two builds

_SCALARS = {"int": ctypes.c_int, "long long": ctypes.c_longlong, "float": ctypes.c_float,
            "unsigned long long": ctypes.c_ulonglong}


class ConvDesc(ctypes.Structure):
    """crd_conv_desc"""
    _fields_ = [(n, ctypes.c_int) for n in (
        "B", "H", "W", "Cin", "ldx", "Ho", "Wo", "Cout", "ldy", "KH", "KW", "stride", "pad",
        "transposed", "in_dtype", "out_dtype", "act", "accumulate", "out_nchw", "w_tap_stride", "w_koff")]


def parse_header(path: str = HEADER):
    """-> {name: (restype, [argtypes], [argnames])} for every CRD_API declaration."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"CRD_API\s+([\w\s]+?)\s*(crd_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3)
        argtypes, argnames = [], []
        if args.strip() != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a or a.startswith("crd_stream_t"):
                    argtypes.append(ctypes.c_void_p)
                else:
                    ty = a.rsplit(" ", 1)[0].replace("const ", "").strip()
                    argtypes.append(_SCALARS[ty])
                argnames.append(a.replace("*", " ").split()[-1])
        protos[name] = (_SCALARS[ret], argtypes, argnames)
    return protos


PROTOS = parse_header()
_lib = None


class KernelError(RuntimeError):
    pass


def load():
    """Load the library (raises if it has not been built: `python -m camradepth_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KernelError(f"{LIB_PATH} is missing: build it with `python -m camradepth_b200.build` "
                          "(there is no CPU fallback by design)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (ret, argtypes, _) in PROTOS.items():
        fn = getattr(lib, name)          # AttributeError if the header declares a missing symbol
        fn.restype = ret
        fn.argtypes = argtypes
    _lib = lib
    return lib


class _Calls:
    """`K.crd_xxx(*args)`: call and raise on a non-zero status."""

    def __getattr__(self, name):
        lib = load()
        fn = getattr(lib, name)
        if PROTOS[name][0] is not ctypes.c_int:
            setattr(self, name, fn)
            return fn

        def call(*args, _fn=fn, _name=name):
            r = _fn(*args)
            if r != 0:
                raise KernelError(f"{_name} failed with status {r}")
        setattr(self, name, call)
        return call


K = _Calls()
