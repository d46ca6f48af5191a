"""TEST INFRASTRUCTURE ONLY -- ctypes loaders for the CPU oracle and the compiled reference.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package. The product (jampack_b200/) never does.

  port()  -> oracle/_build/libjporacle.so : our C restatement (bwt_oracle.c) + generators (gen.c)
  ref()   -> oracle/_ref/libjamref.so     : the UNMODIFIED reference BWT stage (bwt.cpp + divsufsort.cpp
             + format.cpp + sys_detect.cpp of /root/reference) behind oracle/ref_wrap.cpp, or None if it
             was never built (it is built in the authoring container and travels to the GPU box).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
UNITS = 120  # format.hpp:26
TRAILER = UNITS * 4

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)


def _ptr(a, t=_u8p):
    return a.ctypes.data_as(t)


def build(reference_dir="/root/reference"):
    """Compile the oracle (always) and oracle/_ref (only where the reference sources exist)."""
    targets = ["oracle"]
    if os.path.isfile(os.path.join(reference_dir, "bwt.cpp")):
        targets.append("ref")
    subprocess.run(["make", "-s", "-C", HERE, f"REF={reference_dir}"] + targets, check=True)


_port = None
_ref = None


def port():
    global _port
    if _port is None:
        path = os.path.join(HERE, "_build", "libjporacle.so")
        if not os.path.isfile(path):
            build()
        L = C.CDLL(path)
        L.jpo_bwt_forward.argtypes = [_u8p, C.c_int32, _u8p, _i32p]
        L.jpo_bwt_forward.restype = C.c_int
        L.jpo_bwt_inverse.argtypes = [_u8p, C.c_int32, _u8p, _i32p, C.c_int]
        L.jpo_bwt_inverse.restype = C.c_int
        L.jpo_build_map.argtypes = [_u8p, C.c_int32, C.c_int32, _i32p, _i32p]
        L.jpo_build_map.restype = C.c_int
        L.jpo_suffix_array_export.argtypes = [_u8p, _i32p, C.c_int32]
        L.jpo_suffix_array_export.restype = C.c_int
        L.jpo_check_suffix_array.argtypes = [_u8p, _i32p, C.c_int32]
        L.jpo_check_suffix_array.restype = C.c_int
        L.jpo_src_rle0.argtypes = [_u8p, C.c_int32, _i32p, C.POINTER(C.c_uint16), _i32p]
        L.jpo_src_rle0.restype = C.c_int
        _port = L
    return _port


def ref():
    global _ref
    if _ref is None:
        path = os.path.join(HERE, "_ref", "libjamref.so")
        if not os.path.isfile(path):
            return None
        L = C.CDLL(path)
        L.ref_bwt_forward.argtypes = [_u8p, C.c_int, _u8p, _i32p]
        L.ref_bwt_forward.restype = C.c_int
        L.ref_bwt_inverse.argtypes = [_u8p, C.c_int, _u8p, _i32p, C.c_int, _i32p]
        L.ref_bwt_inverse.restype = C.c_int
        pp = C.POINTER(C.c_void_p)
        L.ref_bwt_forward_batch.argtypes = [pp, _i32p, pp, C.c_int, C.c_int]
        L.ref_bwt_forward_batch.restype = C.c_double
        L.ref_bwt_inverse_batch.argtypes = [pp, _i32p, pp, C.c_int, C.c_int, C.c_int]
        L.ref_bwt_inverse_batch.restype = C.c_double
        L.ref_core_count.restype = C.c_int
        L.ref_divsufsort.argtypes = [_u8p, _i32p, C.c_int]
        L.ref_divsufsort.restype = C.c_int
        if hasattr(L, "ref_src_rle0"):
            L.ref_src_rle0.argtypes = [_u8p, C.c_int, _i32p, C.POINTER(C.c_uint16), _i32p]
            L.ref_src_rle0.restype = C.c_int
        _ref = L
    return _ref


# ---- generators live in synth/ (shared with bench.py); re-exported for the tests ------------------------
def gen(kind, n, seed=0):
    import synth
    return synth.gen(kind, n, seed)


def fnv(a):
    import synth
    return synth.fnv(a)


# ---- stage calls; `impl` is "port" or "ref" -------------------------------------------------------
def forward(T, impl="port", prefill=0):
    """-> np.uint8[len+480]; the trailer keeps `prefill` bytes when nlen == 0 (bwt.cpp:35)."""
    T = np.ascontiguousarray(T, dtype=np.uint8)
    out = np.full(T.size + TRAILER, prefill, dtype=np.uint8)
    ol = C.c_int32(0)
    if impl == "ref":
        rc = ref().ref_bwt_forward(_ptr(T), T.size, _ptr(out), C.byref(ol))
    else:
        rc = port().jpo_bwt_forward(_ptr(T), T.size, _ptr(out), C.byref(ol))
    if rc != 0:
        raise RuntimeError(f"oracle forward rc={rc}")
    assert ol.value == T.size + TRAILER
    return out


def inverse(B, impl="port", units=120, threads=1):
    B = np.ascontiguousarray(B, dtype=np.uint8)
    out = np.zeros(max(B.size - TRAILER, 0), dtype=np.uint8)
    ol = C.c_int32(0)
    if impl == "ref":
        after = C.c_int32(0)
        rc = ref().ref_bwt_inverse(_ptr(B), B.size, _ptr(out), C.byref(ol), threads, C.byref(after))
        assert after.value == B.size - TRAILER  # bwt.cpp:77 mutates *Input.size
    else:
        rc = port().jpo_bwt_inverse(_ptr(B), B.size, _ptr(out), C.byref(ol), units)
    if rc != 0:
        raise RuntimeError(f"oracle inverse rc={rc}")
    assert ol.value == B.size - TRAILER
    return out


def indices(B):
    """The 120 sampled primary indices stored after the raw tail (bwt.cpp:57-61)."""
    B = np.ascontiguousarray(B, dtype=np.uint8)
    return np.frombuffer(B[B.size - TRAILER:].tobytes(), dtype="<i4").copy()


def build_map(B, nlen, idx):
    B = np.ascontiguousarray(B, dtype=np.uint8)
    M = np.empty(nlen, dtype=np.int32)
    Ct = np.empty(257, dtype=np.int32)
    port().jpo_build_map(_ptr(B), nlen, idx, _ptr(M, _i32p), _ptr(Ct, _i32p))
    return M, Ct


def suffix_array(T, impl="port"):
    """impl="ref": the reference's divsufsort() itself (divsufsort.cpp:1721)."""
    T = np.ascontiguousarray(T, dtype=np.uint8)
    SA = np.empty(T.size, dtype=np.int32)
    if impl == "ref":
        rc = ref().ref_divsufsort(_ptr(T), _ptr(SA, _i32p), T.size)
    else:
        rc = port().jpo_suffix_array_export(_ptr(T), _ptr(SA, _i32p), T.size)
    assert rc == 0
    return SA


def src_rle0(block, impl="port"):
    """Sorted rank coding + RLE0 per 1 MiB chunk (ans.cpp:149-160) -> (freq int32[chunks, 256], list of uint16 arrays).
    impl="ref": the reference's own Postcoder::Encode / RLE::encode (rank.cpp, rle.cpp compiled into oracle/_ref)."""
    block = np.ascontiguousarray(block, dtype=np.uint8)
    n = block.size
    nchunk = (n + (1 << 20) - 1) >> 20
    freq = np.zeros((max(nchunk, 1), 256), dtype=np.int32)
    rle = np.zeros(max(nchunk, 1) << 20, dtype=np.uint16)
    rlen = np.zeros(max(nchunk, 1), dtype=np.int32)
    fn = ref().ref_src_rle0 if impl == "ref" else port().jpo_src_rle0
    got = fn(_ptr(block), n, _ptr(freq, _i32p), rle.ctypes.data_as(C.POINTER(C.c_uint16)), _ptr(rlen, _i32p))
    assert got == nchunk
    return freq[:nchunk], [rle[(k << 20): (k << 20) + int(rlen[k])].copy() for k in range(nchunk)]
