"""Synthetic inputs for tests and bench.py: splitmix64-driven `uniform`, `markov2`,
`repetitive`, `alla` and the two known-answer formulas, plus FNV-1a-64. C for speed (64-256 MiB blocks),
built on first use with gcc into synth/_build/. Not part of the product path, not part of the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(HERE, "_build", "libjpsynth.so")
_u8p = C.POINTER(C.c_uint8)
_lib = None


def build(force=False):
    src = os.path.join(HERE, "gen.c")
    if force or not os.path.isfile(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_LIB), exist_ok=True)
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-Wall", "-o", _LIB, src], check=True)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        for name in ("uniform", "markov2", "repetitive"):
            f = getattr(L, "jps_gen_" + name)
            f.argtypes = [_u8p, C.c_int64, C.c_uint64]
            f.restype = None
        for name in ("alla", "kat_quadratic", "kat_extremes"):
            f = getattr(L, "jps_gen_" + name)
            f.argtypes = [_u8p, C.c_int64]
            f.restype = None
        L.jps_fnv1a64.argtypes = [_u8p, C.c_int64]
        L.jps_fnv1a64.restype = C.c_uint64
        _lib = L
    return _lib


def gen(kind, n, seed=0, out=None):
    """kind in uniform|markov2|repetitive|alla|kat_quadratic|kat_extremes -> np.uint8[n]"""
    n = int(n)
    T = np.empty(n, dtype=np.uint8) if out is None else out[:n]
    if n == 0:
        return T
    L = lib()
    p = T.ctypes.data_as(_u8p)
    if kind in ("uniform", "markov2", "repetitive"):
        getattr(L, "jps_gen_" + kind)(p, n, int(seed))
    else:
        getattr(L, "jps_gen_" + kind)(p, n)
    return T


def fnv(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return int(lib().jps_fnv1a64(a.ctypes.data_as(_u8p), a.size))
