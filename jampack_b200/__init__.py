"""jampack_b200 -- B200-native (sm_100a) BWT stage for Jampack, behind the C-ABI of include/jp_bwt.h.

The product is `libjpbwt.so` (hand-written CUDA + a thin C-ABI). This module is the Python host mirror
of the reference's stage interface (reference bwt.hpp:13-18, format.hpp:37-54):

    Bwt().ForwardBwt(Input, Output)            # Buffer(block: np.uint8[cap], size: [int])
    Bwt().InverseBwt(Input, Output, Options)

plus array-level helpers (`forward`, `inverse`) and device-resident entry points that take torch CUDA
tensors. There is no CPU path: everything raises `BwtError` if the CUDA library or a device is missing.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libjpbwt.so")
UNITS = 120            # format.hpp:26
TRAILER = UNITS * 4    # bwt.cpp:27
MAX_ROUNDS = 40

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)


class BwtError(RuntimeError):
    def __init__(self, rc, what, detail=""):
        super().__init__(f"{what} (rc={rc}){': ' + detail if detail else ''}")
        self.rc = rc


class Stats(C.Structure):
    _fields_ = [("direction", C.c_int32), ("len", C.c_int32), ("nlen", C.c_int32), ("device", C.c_int32),
                ("kernel_launches", C.c_int32), ("rounds", C.c_int32), ("symbol_bits", C.c_int32),
                ("initial_depth", C.c_int32), ("subchains", C.c_int32), ("subchain_spacing", C.c_int32),
                ("device_bytes", C.c_uint64), ("random_sectors", C.c_uint64),
                ("ms_total", C.c_float), ("ms_h2d", C.c_float), ("ms_d2h", C.c_float),
                ("ms_phase", C.c_float * 8), ("active_fraction", C.c_float * MAX_ROUNDS), ("stream_chunks", C.c_int32),
                ("large_fraction", C.c_float), ("radix_tiles", C.c_int32), ("bypass_suffixes", C.c_int32),
                ("bypass_runs", C.c_int32), ("period", C.c_int32)]

    def asdict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k not in ("ms_phase", "active_fraction")}
        d["ms_phase"] = [round(float(x), 4) for x in self.ms_phase]
        d["active_fraction"] = [round(float(x), 6) for x in self.active_fraction[: max(self.rounds, 0)]]
        return d


_lib = None
# numpy blocks come and go: the array-level helpers give the page-lock back after every call unless the caller says the
# blocks are long-lived (keep_host_blocks_locked(True): what a C++ host with persistent blocks gets by default)
_release_after_call = True


def keep_host_blocks_locked(keep):
    global _release_after_call
    _release_after_call = not keep


def lib():
    """Loads libjpbwt.so; fails loudly (no fallback) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise BwtError(-2, f"{LIB_PATH} is missing: build it with `python -m jampack_b200.build` "
                               "(this stage has no CPU path)")
        L = C.CDLL(LIB_PATH)
        L.jp_bwt_forward.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, _i32p]
        L.jp_bwt_inverse.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, _i32p]
        L.jp_bwt_forward_device.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int, C.c_void_p]
        L.jp_bwt_inverse_device.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int, C.c_void_p]
        L.jp_bwt_inverse_device_consume.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int, C.c_void_p]
        L.jp_bwt_set_devices.argtypes = [C.POINTER(C.c_int), C.c_int]
        L.jp_bwt_device_count.argtypes = []
        L.jp_bwt_host_alloc.argtypes = [C.c_uint64]
        L.jp_bwt_host_alloc.restype = C.c_void_p
        L.jp_bwt_host_free.argtypes = [C.c_void_p]
        L.jp_bwt_host_free.restype = None
        L.jp_bwt_host_release.argtypes = [C.c_void_p]
        L.jp_bwt_host_release.restype = None
        L.jp_bwt_last_stats.argtypes = [C.POINTER(Stats)]
        L.jp_bwt_strerror.argtypes = [C.c_int]
        L.jp_bwt_strerror.restype = C.c_char_p
        L.jp_bwt_last_error_detail.restype = C.c_char_p
        L.jp_bwt_version.restype = C.c_char_p
        L.jp_bwt_debug_lf.argtypes = [C.c_void_p, C.c_int32, _i32p, _i32p]
        L.jp_bwt_suffix_array.argtypes = [C.c_void_p, C.c_int32, _i32p]
        L.jp_bwt_debug_gather_rate.argtypes = [C.c_uint64, C.c_int32, C.c_int32, C.c_int]
        L.jp_bwt_debug_gather_rate.restype = C.c_double
        L.jp_src_rle0_device.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.jp_src_rle0.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.jp_bwt_debug_copy.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
        L.jp_bwt_warmup_async.argtypes = []
        _lib = L
    return _lib


EXPORTS = ["jp_bwt_forward", "jp_bwt_inverse", "jp_bwt_forward_device", "jp_bwt_inverse_device", "jp_bwt_inverse_device_consume",
           "jp_bwt_set_devices", "jp_bwt_device_count", "jp_bwt_host_alloc", "jp_bwt_host_free", "jp_bwt_host_release",
           "jp_bwt_last_stats", "jp_bwt_strerror", "jp_bwt_last_error_detail", "jp_bwt_version",
           "jp_bwt_debug_lf", "jp_bwt_suffix_array", "jp_bwt_debug_gather_rate", "jp_bwt_debug_copy", "jp_bwt_warmup_async", "jp_src_rle0_device", "jp_src_rle0"]


def _check(rc, what):
    if rc != 0:
        L = lib()
        raise BwtError(rc, f"{what}: {L.jp_bwt_strerror(rc).decode()}", L.jp_bwt_last_error_detail().decode())


def last_stats():
    s = Stats()
    lib().jp_bwt_last_stats(C.byref(s))
    return s


def set_devices(ids):
    arr = (C.c_int * max(len(ids), 1))(*ids)
    _check(lib().jp_bwt_set_devices(arr, len(ids)), "jp_bwt_set_devices")


def device_count():
    return lib().jp_bwt_device_count()


class PinnedBlock:
    """A page-locked host block (jp_bwt_host_alloc) exposed as a numpy uint8 array."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        self.ptr = lib().jp_bwt_host_alloc(self.nbytes)
        if not self.ptr:
            raise BwtError(-4, "jp_bwt_host_alloc failed")
        self.array = np.ctypeslib.as_array(C.cast(self.ptr, _u8p), shape=(max(self.nbytes, 1),))[: self.nbytes]

    def free(self):
        if self.ptr:
            self.array = None
            lib().jp_bwt_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def host_release(array=None):
    """Drops the page-lock the library took on a pageable block (all of them when `array` is None): call it before such
    a block is freed while the library stays loaded."""
    lib().jp_bwt_host_release(None if array is None else array.ctypes.data)


# ---- array-level host entry points -------------------------------------------------------------------
def forward(block, out=None, prefill=0):
    """BWT of a host block. Returns np.uint8[len + 480] (BWT bytes, raw tail, 120 int32 indices).
    For len < 120 the trailer bytes are left as found in `out` (or `prefill`), like the reference (bwt.cpp:35)."""
    block = np.ascontiguousarray(block, dtype=np.uint8)
    n = block.size
    if out is None:
        out = np.full(n + TRAILER, prefill, dtype=np.uint8)
    assert out.dtype == np.uint8 and out.size >= n + TRAILER and out.flags.c_contiguous
    ol = C.c_int32(0)
    try:
        _check(lib().jp_bwt_forward(block.ctypes.data, n, out.ctypes.data, C.byref(ol)), "jp_bwt_forward")
    finally:
        if _release_after_call:
            host_release(block); host_release(out)
    assert ol.value == n + TRAILER
    return out[: n + TRAILER]


def inverse(block, out=None):
    """Inverse of `forward`: np.uint8[len_with_trailer] -> np.uint8[len_with_trailer - 480]."""
    block = np.ascontiguousarray(block, dtype=np.uint8)
    n = block.size
    if out is None:
        out = np.zeros(max(n - TRAILER, 0), dtype=np.uint8)
    ol = C.c_int32(0)
    try:
        _check(lib().jp_bwt_inverse(block.ctypes.data, n, out.ctypes.data, C.byref(ol)), "jp_bwt_inverse")
    finally:
        if _release_after_call:
            host_release(block); host_release(out)
    assert ol.value == n - TRAILER
    return out[: n - TRAILER]


# ---- device-resident entry points (torch CUDA uint8 tensors) ------------------------------------------
def _stream_of(t):
    """The torch stream the caller's tensors are ordered on. The legacy default stream has handle 0, which the C-ABI
    reads as "use the context's own (non-blocking) stream": pending torch work on it is drained first, or the stage
    could start before the caller's fill/clone of its blocks has finished."""
    import torch
    st = torch.cuda.current_stream(t.device)
    if st.cuda_stream == 0:
        st.synchronize()
    return st.cuda_stream


def forward_device(d_in, d_out=None):
    import torch
    assert d_in.is_cuda and d_in.dtype == torch.uint8 and d_in.is_contiguous()
    n = d_in.numel()
    if d_out is None:
        d_out = torch.zeros(n + TRAILER, dtype=torch.uint8, device=d_in.device)
    assert d_out.numel() >= n + TRAILER
    _check(lib().jp_bwt_forward_device(d_in.data_ptr(), n, d_out.data_ptr(), d_in.device.index or 0, _stream_of(d_in)),
           "jp_bwt_forward_device")
    return d_out


def inverse_device(d_in, d_out=None, consume=False):
    """consume=True lets the stage overwrite d_in with scratch data (stays within 6N of device memory)."""
    import torch
    assert d_in.is_cuda and d_in.dtype == torch.uint8 and d_in.is_contiguous()
    n = d_in.numel()
    assert n >= TRAILER, "an inverse input carries the 480-byte trailer"
    if d_out is None:
        d_out = torch.zeros(max(n - TRAILER, 1), dtype=torch.uint8, device=d_in.device)
    assert d_out.is_cuda and d_out.dtype == torch.uint8 and d_out.is_contiguous() and d_out.numel() >= n - TRAILER
    fn = lib().jp_bwt_inverse_device_consume if consume else lib().jp_bwt_inverse_device
    _check(fn(d_in.data_ptr(), n, d_out.data_ptr(), d_in.device.index or 0, _stream_of(d_in)), "jp_bwt_inverse_device")
    return d_out


# ---- second stage, first half: sorted rank coding + RLE0 per 1 MiB chunk (reference rank.cpp:45-90, rle.cpp:22-47) ----
ANS_CHUNK = 1 << 20        # ans.hpp:33 StackSize


def src_rle0(block):
    """Host block -> (freq int32[chunks, 256], rle: list of uint16 arrays, one per chunk)."""
    block = np.ascontiguousarray(block, dtype=np.uint8)
    n = block.size
    nchunk = (n + ANS_CHUNK - 1) // ANS_CHUNK
    freq = np.zeros((max(nchunk, 1), 256), dtype=np.int32)
    rle = np.zeros(max(n, 1), dtype=np.uint16)
    rlen = np.zeros(max(nchunk, 1), dtype=np.int32)
    _check(lib().jp_src_rle0(block.ctypes.data, n, freq.ctypes.data, rle.ctypes.data, rlen.ctypes.data), "jp_src_rle0")
    return freq[:nchunk], [rle[k * ANS_CHUNK: k * ANS_CHUNK + int(rlen[k])].copy() for k in range(nchunk)]


def src_rle0_device(d_in):
    """The same on a block that is resident in HBM (e.g. what forward_device returned): -> (freq, rle, rlen) CUDA tensors;
    chunk k's symbols are rle[k * ANS_CHUNK : k * ANS_CHUNK + rlen[k]]."""
    import torch
    assert d_in.is_cuda and d_in.dtype == torch.uint8 and d_in.is_contiguous()
    n = d_in.numel()
    nchunk = (n + ANS_CHUNK - 1) // ANS_CHUNK
    freq = torch.zeros((max(nchunk, 1), 256), dtype=torch.int32, device=d_in.device)
    rle = torch.zeros(max(n, 1), dtype=torch.int16, device=d_in.device)
    rlen = torch.zeros(max(nchunk, 1), dtype=torch.int32, device=d_in.device)
    _check(lib().jp_src_rle0_device(d_in.data_ptr(), n, freq.data_ptr(), rle.data_ptr(), rlen.data_ptr(), d_in.device.index or 0, _stream_of(d_in)),
           "jp_src_rle0_device")
    return freq[:nchunk], rle, rlen[:nchunk]


# ---- the reference's stage interface, mirrored ---------------------------------------------------------
class Buffer:
    """format.hpp:37-41 -- `block` is a uint8 array with spare capacity, `size` a one-element list (Index*)."""

    def __init__(self, block, size=None):
        self.block = block
        self.size = [int(block.size if size is None else size)]


class Options:
    """format.hpp:46-54. Only carried for signature parity: the GPU inverse runs all 120 units at once."""

    def __init__(self, BlockSize=8 << 20, MatchFinder=0, Threads=1, Filters=1, Gpu=True, Multiblock=True):
        self.BlockSize, self.MatchFinder, self.Threads = BlockSize, MatchFinder, Threads
        self.Filters, self.Gpu, self.Multiblock = Filters, Gpu, Multiblock


def Error(msg):
    """format.cpp:6-10 prints and exits; a library raises instead."""
    raise BwtError(-1, msg)


class Bwt:
    """BlockSort::Bwt (bwt.hpp:13-18)."""

    def ForwardBwt(self, Input, Output):
        n = Input.size[0]
        if Output.block.size < n + TRAILER:
            Error("Bwt :: output block too small")
        ol = C.c_int32(0)
        rc = lib().jp_bwt_forward(Input.block.ctypes.data, n, Output.block.ctypes.data, C.byref(ol))
        Output.size[0] = n + TRAILER                     # bwt.cpp:27
        if rc != 0:
            Error(lib().jp_bwt_strerror(rc).decode())

    def InverseBwt(self, Input, Output, Opt=None):
        n = Input.size[0]
        ol = C.c_int32(0)
        rc = lib().jp_bwt_inverse(Input.block.ctypes.data, n, Output.block.ctypes.data, C.byref(ol))
        Input.size[0] = n - TRAILER                      # bwt.cpp:77 mutates *Input.size
        Output.size[0] = Input.size[0]                   # bwt.cpp:78
        if rc != 0:
            Error(lib().jp_bwt_strerror(rc).decode())


# ---- test hooks ------------------------------------------------------------------------------------------
def debug_lf(bwt_bytes):
    b = np.ascontiguousarray(bwt_bytes, dtype=np.uint8)
    lf = np.empty(b.size, dtype=np.int32)
    ct = np.empty(257, dtype=np.int32)
    _check(lib().jp_bwt_debug_lf(b.ctypes.data, b.size, lf.ctypes.data_as(_i32p), ct.ctypes.data_as(_i32p)), "jp_bwt_debug_lf")
    return lf, ct


def suffix_array(text):
    """The forward transform's suffix sorter on its own: divsufsort(T, SA, n) semantics (reference divsufsort.hpp:37-45)."""
    t = np.ascontiguousarray(text, dtype=np.uint8)
    sa = np.empty(t.size, dtype=np.int32)
    _check(lib().jp_bwt_suffix_array(t.ctypes.data, t.size, sa.ctypes.data_as(_i32p)), "jp_bwt_suffix_array")
    return sa


debug_suffix_array = suffix_array


def debug_copy(src, dst):
    """Only the copies of a stage call: host `src` -> device, device -> host `dst` (bench.py's copy ceiling)."""
    _check(lib().jp_bwt_debug_copy(src.ctypes.data, src.size, dst.ctypes.data, dst.size), "jp_bwt_debug_copy")


def debug_gather_rate(table_bytes, chains, steps, dependent=True):
    r = lib().jp_bwt_debug_gather_rate(int(table_bytes), int(chains), int(steps), 1 if dependent else 0)
    if r < 0:
        _check(int(r), "jp_bwt_debug_gather_rate")
    return r
