"""Extended differential run (too long for the test suite): random blocks of 0.5-12 MiB with random alphabets and
structure through forward and inverse against the compiled reference.  python tools/stress.py [cases] [seed]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import jampack_b200 as jp  # noqa: E402
import oracle  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
impl = "ref" if oracle.ref() is not None else "port"
bad = 0
t0 = time.time()
for c in range(cases):
    n = int(rng.integers(1 << 19, 12 << 20))
    kind = int(rng.integers(0, 7))
    sigma = int(rng.choice([2, 3, 5, 17, 64, 200, 256]))
    if kind == 0:
        T = rng.integers(0, sigma, n).astype(np.uint8)
    elif kind == 1:
        T = oracle.gen("markov2", n, int(rng.integers(1, 1 << 30)))
    elif kind == 2:                                   # copy-paste text: long repeats at random offsets
        T = oracle.gen("markov2", n, int(rng.integers(1, 1 << 30)))
        for _ in range(int(rng.integers(1, 30))):
            L = int(rng.integers(100, n // 4)); a = int(rng.integers(0, n - L)); b = int(rng.integers(0, n - L))
            T[b:b + L] = T[a:a + L].copy()
    elif kind == 3:                                   # runs
        T = np.repeat(rng.integers(0, sigma, n // 50 + 1).astype(np.uint8), rng.integers(1, 100, n // 50 + 1))[:n]
        n = T.size
    elif kind == 4:                                   # periodic with defects
        p = int(rng.integers(1, 5000))
        T = np.tile(rng.integers(0, sigma, p).astype(np.uint8), n // p + 1)[:n].copy()
        T[rng.integers(0, n, int(rng.integers(1, 50)))] ^= 1
    elif kind == 5:                                   # zero pages + data
        T = rng.integers(0, sigma, n).astype(np.uint8)
        for _ in range(int(rng.integers(1, 20))):
            a = int(rng.integers(0, n)); T[a:a + int(rng.integers(1, 1 << 18))] = 0
    else:
        T = oracle.gen("repetitive", n, int(rng.integers(1, 1 << 30)))
    want = oracle.forward(T, impl, prefill=1)
    got = jp.forward(T, prefill=1)
    fs = jp.last_stats()
    ok_f = bool((got == want).all())
    back = jp.inverse(want)
    chunks = jp.last_stats().stream_chunks      # > 0: single-walk inverse (JP_BWT_INV_SINGLE=1 forces it on these sizes)
    ok_i = bool((back == T).all())
    print(f"case {c:3d} kind={kind} n={n:9d} sigma={np.unique(T).size:3d} rounds={fs.rounds:2d} fwd={fs.ms_total:8.2f} ms large={fs.large_fraction:.3f} "
          f"forward={'ok' if ok_f else 'MISMATCH'} inverse={'ok' if ok_i else 'MISMATCH'} stream_chunks={chunks}", flush=True)
    bad += (not ok_f) + (not ok_i)
print(f"STRESS {cases} cases, {bad} failures, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
