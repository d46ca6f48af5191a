"""Small cases through every forward shortcut and the second-stage kernels, meant to run under compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_r02.py
Test infrastructure, not product."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import jampack_b200 as jp
import oracle


def periodic(n, p, sig, defects, seed):
    r = np.random.default_rng(seed)
    T = np.tile(r.integers(0, sig, p).astype(np.uint8), n // p + 1)[:n].copy()
    T[r.integers(0, n, defects)] ^= 1
    return T


rng = np.random.default_rng(1)
zero_pages = rng.integers(0, 256, 300000).astype(np.uint8); zero_pages[40000:200000] = 0; zero_pages[250000:260000] = 7
tar_like = oracle.gen("markov2", 200000, 3)
for a in range(0, 200000 - 512, 1024):
    tar_like[a + 300 + (a // 512) % 150: a + 512] = 0
cases = [("all-a", np.full(250000, 97, np.uint8), {}), ("zero pages", zero_pages, {"JP_BWT_FWD_BYPASS": "1"}), ("tar-like", tar_like, {"JP_BWT_FWD_BYPASS": "1", "JP_BWT_FWD_RUNJUMP": "1"}),
         ("repetitive / representatives", oracle.gen("repetitive", 400000, 3), {"JP_BWT_FWD_REDUCED": "1"}), ("period 60 / representatives", periodic(300000, 60, 5, 9, 2), {"JP_BWT_FWD_REDUCED": "1"}),
         ("period 300 / repeat lengths", periodic(200000, 300, 2, 5, 4), {"JP_BWT_FWD_PERIODIC": "1", "JP_BWT_FWD_REDUCED": "0"}),
         ("markov2 staged ranks", oracle.gen("markov2", 300000, 1), {"JP_BWT_ISA_STAGE_MIN": "1000", "JP_BWT_ISA_REGION_LOG2": "12"}),
         ("all-a without bypass (large route, batches)", np.full(200000, 5, np.uint8), {"JP_BWT_FWD_BYPASS": "0", "JP_BWT_FWD_PERIODIC": "0"}),
         ("markov2 coded keys, order 2, packed", oracle.gen("markov2", 300000, 2), {"JP_BWT_FWD_CTXKEYS": "1", "JP_BWT_FWD_PACKED": "1"}),
         ("markov2 coded keys, order 2, pairs", oracle.gen("markov2", 300000 + 77, 2), {"JP_BWT_FWD_CTXKEYS": "1", "JP_BWT_FWD_PACKED": "0"}),
         ("uniform coded keys, order 1, 63 bits", oracle.gen("uniform", 150000, 2), {"JP_BWT_FWD_CTXKEYS": "1", "JP_BWT_FWD_KEYPASSES": "8"}),
         ("dna coded keys, short block", (rng.integers(0, 4, 9000) + 65).astype(np.uint8), {"JP_BWT_FWD_CTXKEYS": "1"})]
keys = ("JP_BWT_FWD_BYPASS", "JP_BWT_FWD_RUNJUMP", "JP_BWT_FWD_REDUCED", "JP_BWT_FWD_PERIODIC", "JP_BWT_ISA_STAGE_MIN", "JP_BWT_ISA_REGION_LOG2",
        "JP_BWT_FWD_CTXKEYS", "JP_BWT_FWD_PACKED", "JP_BWT_FWD_KEYPASSES")
bad = 0
for name, T, env in cases:
    for k in keys:
        os.environ.pop(k, None)
    os.environ.update(env)
    want = oracle.forward(T, "ref" if oracle.ref() is not None else "port")
    got = jp.forward(T)
    st = jp.last_stats()
    ok = bool((got == want).all()) and bool((jp.inverse(got) == T).all())
    f, r = jp.src_rle0(got)
    fw, rw = oracle.src_rle0(want, "port")
    ok2 = bool((f == fw).all()) and all(a.size == b.size and (a == b).all() for a, b in zip(r, rw))
    print(f"{name:45s} forward+inverse {'ok' if ok else 'MISMATCH'} stage2 {'ok' if ok2 else 'MISMATCH'} rounds={st.rounds} bypass={st.bypass_suffixes}/{st.bypass_runs} period={st.period}", flush=True)
    bad += (not ok) + (not ok2)
print("SANITIZE CASES", len(cases), "failures", bad)
sys.exit(1 if bad else 0)
