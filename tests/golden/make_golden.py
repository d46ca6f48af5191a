"""Generates tests/golden/kat.json from the UNMODIFIED reference compiled into oracle/_ref.

Run in the authoring container (needs /root/reference for `make -C oracle ref`):
    python tests/golden/make_golden.py [--big]
The reference ships no golden vectors, so these known-answer values ARE the pin:
every value below is the output of BlockSort::Bwt::ForwardBwt (bwt.cpp:22-65) itself, each one
round-tripped through the reference InverseBwt (bwt.cpp:72-282). `fnv` is FNV-1a-64 as implemented
in oracle/gen.c. --big adds the 64 MiB / 256 MiB blocks of BASELINE.json configs 2, 3 and 5.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np  # noqa: E402
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
MiB = 1 << 20

SMALL = [
    ("KAT-A", "kat_quadratic", 240, 0), ("KAT-B", "kat_quadratic", 250, 0), ("KAT-C", "kat_quadratic", 119, 0),
    ("KAT-D", "alla", 360, 0), ("KAT-E", "kat_extremes", 240, 0),
    ("KAT-F", "kat_quadratic", 120, 0), ("KAT-G", "uniform", 121, 7), ("KAT-H", "markov2", 4093, 9),
    ("KAT-I", "repetitive", 70000, 3), ("KAT-J", "kat_extremes", 1, 0),
]
MEDIUM = [("uniform-1M", "uniform", MiB, 2), ("markov2-1M", "markov2", MiB, 1), ("repetitive-1M", "repetitive", MiB, 3),
          ("alla-1M", "alla", MiB, 0), ("markov2-8M", "markov2", 8 * MiB, 1)]
BIG = [("markov2-64M", "markov2", 64 * MiB, 1), ("uniform-64M", "uniform", 64 * MiB, 2),
       ("repetitive-64M", "repetitive", 64 * MiB, 3), ("alla-64M", "alla", 64 * MiB, 0),
       ("markov2-256M", "markov2", 256 * MiB, 5)]


def record(name, kind, n, seed, full):
    T = oracle.gen(kind, n, seed)
    out = oracle.forward(T, "ref", prefill=0)
    back = oracle.inverse(out, "ref", threads=8)
    assert (back == T).all(), name
    nlen = n - n % 120
    r = {"name": name, "kind": kind, "len": n, "seed": seed, "nlen": nlen,
         "fnv_in": "%016x" % oracle.fnv(T), "fnv_bwt": "%016x" % oracle.fnv(out[:n]),
         "head48": out[:min(48, n)].tobytes().hex()}
    if nlen > 0:
        r["fnv_all"] = "%016x" % oracle.fnv(out)
        r["indices"] = [int(x) for x in oracle.indices(out)]
    if full:
        r["out_hex"] = out[: n + (480 if nlen > 0 else 0)].tobytes().hex()
    return r


def config1():
    """BASELINE.json configs[0]: 8 MiB markov2 block through the reference CLI, defaults, -t >= 2."""
    exe = os.path.join(os.path.dirname(oracle.__file__), "_ref", "Jampack_ref")
    T = oracle.gen("markov2", 8 * MiB, 1)
    with tempfile.TemporaryDirectory() as d:
        src, jam, back = (os.path.join(d, x) for x in ("in.bin", "out.jam", "back.bin"))
        T.tofile(src)
        subprocess.run([exe, "c", src, jam, "-t4"], check=True, stdout=subprocess.DEVNULL)
        subprocess.run([exe, "d", jam, back, "-t4"], check=True, stdout=subprocess.DEVNULL)
        jb = open(jam, "rb").read()
        assert open(back, "rb").read() == T.tobytes()
        return {"input_sha256": hashlib.sha256(T.tobytes()).hexdigest(), "jam_bytes": len(jb),
                "jam_sha256": hashlib.sha256(jb).hexdigest(), "flags": "-t4 (any -t >= 2), defaults otherwise"}


def main():
    assert oracle.ref() is not None, "build oracle/_ref first: make -C oracle ref"
    path = os.path.join(HERE, "kat.json")
    old = json.load(open(path)) if os.path.isfile(path) else {}
    doc = {"source": "reference BlockSort::Bwt::ForwardBwt via oracle/_ref/libjamref.so (bwt.cpp:22-65)",
           "cases": [record(*c, full=True) for c in SMALL] + [record(*c, full=False) for c in MEDIUM]}
    if "--big" in sys.argv:
        doc["big"] = [record(*c, full=False) for c in BIG]
        doc["config1"] = config1()
    else:
        for k in ("big", "config1"):
            if k in old:
                doc[k] = old[k]
    json.dump(doc, open(path, "w"), indent=1)
    print("wrote", path, len(doc["cases"]), "cases", len(doc.get("big", [])), "big")


if __name__ == "__main__":
    main()
