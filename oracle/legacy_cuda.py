"""TEST INFRASTRUCTURE ONLY. Times the reference's own legacy CUDA inverse (oracle/_ref/libjamref_cuda.so, built by
`make -C oracle ref_cuda` from the unmodified sources; see oracle/ref_wrap_cuda.cpp) on one block and prints one JSON
line. bench.py runs this as a child process with a timeout: the legacy path was written for cc <= 6 devices and is
a reported baseline (BASELINE.md plan item 5), so nothing it does may take the benchmark down.

    python -m oracle.legacy_cuda <block-with-trailer.npy> <original-text.npy> [repeats]
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libjamref_cuda.so")


def available():
    return os.path.isfile(LIB)


def main(argv):
    B = np.load(argv[1]); T = np.load(argv[2])
    reps = int(argv[3]) if len(argv) > 3 else 2
    L = C.CDLL(LIB)
    u8p = C.POINTER(C.c_uint8)
    L.ref_bwt_inverse_legacy_cuda.argtypes = [u8p, C.c_int, u8p, C.POINTER(C.c_int)]
    L.ref_bwt_inverse_legacy_cuda.restype = C.c_double
    out = np.zeros(B.size, dtype=np.uint8)
    olen = C.c_int(0)
    best = None
    for _ in range(reps):
        src = B.copy()                                  # (InverseBwt does not write its input, but stay on the safe side)
        s = L.ref_bwt_inverse_legacy_cuda(src.ctypes.data_as(u8p), int(src.size), out.ctypes.data_as(u8p), C.byref(olen))
        if s < 0:
            print(json.dumps({"unavailable": f"no usable CUDA device for the legacy path (code {s})"})); return 0
        best = s if best is None else min(best, s)
    ok = bool(olen.value == T.size and (out[: T.size] == T).all())
    print(json.dumps({"value": round(T.size / best / 1e6, 2), "unit": "MB/s", "seconds": round(best, 4), "output_matches": ok,
                      "kind": "reference, legacy CUDA inverse (bwt.cpp:8-19, :186-240; 120 device threads, Map built on the host, 6N copied in)",
                      "sample": f"1 x {T.size >> 20} MiB inverse block, best of {reps}, wall clock of InverseBwt with Options::Gpu"}))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
