"""Generates tests/golden/stage2.json: known-answer values of the second stage's first half (sorted rank coding + RLE0 per
1 MiB chunk, reference ans.cpp:149-160 -> rank.cpp:45-90, rle.cpp:22-47) from the UNMODIFIED reference compiled into
oracle/_ref (run in the authoring container, where /root/reference exists):   python tests/golden/make_golden_stage2.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import oracle  # noqa: E402

CASES = [("kat_quadratic", 240, 0), ("alla", 360, 0), ("markov2", 4093, 9), ("uniform", 70000, 2), ("markov2", 300000, 1),
         ("repetitive", (1 << 20) + 4567, 3), ("markov2", 3 * (1 << 20), 7), ("alla", (1 << 20) + 120, 0)]


def main():
    assert oracle.ref() is not None and hasattr(oracle.ref(), "ref_src_rle0"), "needs oracle/_ref built from /root/reference"
    out = []
    for kind, n, seed in CASES:
        T = oracle.gen(kind, n, seed)
        B = oracle.forward(T, "ref")                      # the stage input is the BWT block incl. its trailer (jampack.cpp:40-41)
        freq, rle = oracle.src_rle0(B, "ref")
        c = {"kind": kind, "len": n, "seed": seed, "block_bytes": int(B.size), "chunks": len(rle), "rlen": [int(r.size) for r in rle],
             "fnv_freq": "%016x" % oracle.fnv(np.ascontiguousarray(freq).view(np.uint8).ravel()),
             "fnv_rle": ["%016x" % oracle.fnv(np.ascontiguousarray(r).view(np.uint8)) for r in rle]}
        if B.size <= 1000:
            c["freq_nonzero"] = {str(i): int(v) for i, v in enumerate(freq[0]) if v}
            c["rle"] = [int(x) for x in rle[0]]
        out.append(c)
    with open(os.path.join(HERE, "stage2.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden_stage2.py", "source": "reference rank.cpp + rle.cpp via oracle/_ref/libjamref.so", "cases": out}, f, indent=1)
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()
