"""TEST INFRASTRUCTURE ONLY: builds and loads the SIMT-emulated copy of the product's kernels (see simt_runtime.hpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libjpemu.so")
_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_lib = None


def _stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    csrc = os.path.join(os.path.dirname(os.path.dirname(HERE)), "jampack_b200", "csrc")
    deps = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh"))]
    deps += [os.path.join(HERE, f) for f in ("simt_runtime.hpp", "simt_runtime.cpp", "harness.cpp", "gen.py")]
    deps.append(os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "jp_bwt.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False):
    if not force and not _stale():
        return LIB
    from . import gen
    gen.main(["common.cuh", "bwt_internal.cuh", "radix_sort.cuh", "inv_stream.cuh", "bwt_inverse.cu", "bwt_forward.cu", "src_rle0.cu"])
    cmd = ["g++", "-std=c++17", "-O1", "-w", "-fPIC", "-shared", "-I", HERE, "-I", OUT, "-o", LIB,
           os.path.join(OUT, "bwt_inverse.cpp"), os.path.join(OUT, "bwt_forward.cpp"), os.path.join(OUT, "src_rle0.cpp"),
           os.path.join(HERE, "harness.cpp"), os.path.join(HERE, "simt_runtime.cpp")]
    env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
    subprocess.run(cmd, check=True, env=env)
    return LIB


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.emu_inverse.argtypes = [_u8p, C.c_int32, _u8p, C.c_int, _i32p, _i32p]
        L.emu_forward.argtypes = [_u8p, C.c_int32, _u8p, _i32p, _i32p]
        L.emu_suffix_array.argtypes = [_u8p, C.c_int32, _i32p]
        L.emu_src_rle0.argtypes = [_u8p, C.c_int32, _i32p, C.POINTER(C.c_uint16), _i32p]
        _lib = L
    return _lib


def inverse(B, consume=False):
    """-> (rc, text, stream_chunks, kernel launches) of the emulated jp::inverse_device"""
    B = np.ascontiguousarray(B, dtype=np.uint8)
    out = np.zeros(max(B.size - 480, 0), dtype=np.uint8)
    ch, la = C.c_int32(0), C.c_int32(0)
    rc = lib().emu_inverse(B.ctypes.data_as(_u8p), B.size, out.ctypes.data_as(_u8p), int(consume), C.byref(ch), C.byref(la))
    return rc, out, ch.value, la.value


def forward(T, prefill=0x5C):
    """-> (rc, BWT || tail || trailer, rounds, kernel launches) of the emulated jp::forward_device. The emulated device
    output block is pre-filled with 0x5C, so blocks under 120 bytes leave that in the trailer."""
    T = np.ascontiguousarray(T, dtype=np.uint8)
    out = np.zeros(T.size + 480, dtype=np.uint8)
    r, la = C.c_int32(0), C.c_int32(0)
    rc = lib().emu_forward(T.ctypes.data_as(_u8p), T.size, out.ctypes.data_as(_u8p), C.byref(r), C.byref(la))
    return rc, out, r.value, la.value


def suffix_array(T):
    """-> (rc, SA) of the emulated suffix sorter (jp::debug_suffix_array)"""
    T = np.ascontiguousarray(T, dtype=np.uint8)
    sa = np.zeros(max(T.size, 1), dtype=np.int32)
    rc = lib().emu_suffix_array(T.ctypes.data_as(_u8p), T.size, sa.ctypes.data_as(_i32p))
    return rc, sa[: T.size]


def src_rle0(block):
    """-> (rc, freq int32[chunks, 256], list of uint16 arrays) of the emulated jp::src_rle0_device"""
    block = np.ascontiguousarray(block, dtype=np.uint8)
    n = block.size
    nchunk = (n + (1 << 20) - 1) >> 20
    freq = np.zeros((max(nchunk, 1), 256), dtype=np.int32)
    rle = np.zeros(max(n, 1), dtype=np.uint16)
    rlen = np.zeros(max(nchunk, 1), dtype=np.int32)
    rc = lib().emu_src_rle0(block.ctypes.data_as(_u8p), n, freq.ctypes.data_as(_i32p), rle.ctypes.data_as(C.POINTER(C.c_uint16)), rlen.ctypes.data_as(_i32p))
    return rc, freq[:nchunk], [rle[(k << 20): (k << 20) + int(rlen[k])].copy() for k in range(nchunk)]
