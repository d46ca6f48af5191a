"""Differential fuzzing of the product's kernel code under the SIMT emulator (tests/simt) against the oracle: random
structured blocks (noise over 1..256 symbols, runs, periodic data with defects, near-constant blocks, order-2 text,
copy-paste text, plateaus) through the forward (run bypass forced off and on, rank staging forced on) and through both
inverse paths (two-pass, single-walk). CPU only, test infrastructure.
    python tools/emu_fuzz.py [cases] [seed]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import oracle  # noqa: E402
import simt  # noqa: E402

SWITCHES = ("JP_BWT_FWD_BYPASS", "JP_BWT_FWD_PERIODIC", "JP_BWT_FWD_RUNJUMP", "JP_BWT_FWD_REDUCED", "JP_BWT_ISA_STAGE_MIN", "JP_BWT_ISA_REGION_LOG2", "JP_BWT_INV_SINGLE", "JP_BWT_INV_LOG2M", "JP_BWT_INV_WBLOCKS_PER_SM",
            "JP_BWT_FWD_CTXKEYS", "JP_BWT_FWD_PACKED", "JP_BWT_FWD_KEYPASSES", "JP_BWT_FWD_EMIT_REGION_LOG2", "JP_BWT_INV_RANK_PLAN")


def block(rng, n):
    kind = int(rng.integers(0, 8)); sig = int(rng.choice([1, 2, 3, 5, 17, 200, 256]))
    if kind == 0:
        T = rng.integers(0, sig, n).astype(np.uint8)
    elif kind == 1:
        T = np.resize(np.repeat(rng.integers(0, sig, n // 5 + 1).astype(np.uint8), rng.integers(1, int(rng.integers(2, 120)), n // 5 + 1)), n)
    elif kind == 2:
        T = rng.integers(0, sig, n).astype(np.uint8)
        for _ in range(int(rng.integers(1, 8))):
            a = int(rng.integers(0, n)); T[a:a + int(rng.integers(1, n // 2))] = rng.integers(0, 256)
    elif kind == 3:
        p = int(rng.integers(1, 60)); T = np.tile(rng.integers(0, sig, p).astype(np.uint8), n // p + 1)[:n].copy(); T[rng.integers(0, n, 3)] ^= 1
    elif kind == 4:
        T = np.full(n, int(rng.integers(0, 256)), np.uint8); T[rng.integers(0, n, int(rng.integers(0, 4)))] = rng.integers(0, 256)
    elif kind == 5:
        T = oracle.gen("markov2", n, int(rng.integers(1, 1 << 30)))
    elif kind == 6:
        T = oracle.gen("markov2", n, int(rng.integers(1, 1 << 30)))
        for _ in range(int(rng.integers(1, 10))):
            L = int(rng.integers(10, n // 3)); a = int(rng.integers(0, n - L)); b = int(rng.integers(0, n - L)); T[b:b + L] = T[a:a + L].copy()
    else:
        T = np.zeros(n, np.uint8)
        for i, cut in enumerate(np.sort(rng.integers(0, n, int(rng.integers(1, 6))))):
            T[cut:] = (i + 1) % 3
    return np.ascontiguousarray(T[:n])


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    simt.build()
    saved = {k: os.environ.get(k) for k in SWITCHES}
    fails, t0 = 0, time.time()
    for c in range(cases):
        for k in SWITCHES:
            os.environ.pop(k, None)
        big = c % 3 == 0
        T = block(rng, int(rng.integers(66000, 150000)) if big else int(rng.integers(121, 7000)))
        want = oracle.forward(T, "port", prefill=0x5C)
        if not big:                                        # the forward is slow under emulation: small blocks only
            for fv in ({"JP_BWT_FWD_BYPASS": "0", "JP_BWT_FWD_PERIODIC": "0"}, {"JP_BWT_FWD_BYPASS": "1", "JP_BWT_FWD_RUNJUMP": "1"}, {"JP_BWT_FWD_PERIODIC": "1"},
                       {"JP_BWT_FWD_BYPASS": "1", "JP_BWT_FWD_PERIODIC": "1", "JP_BWT_ISA_STAGE_MIN": "100", "JP_BWT_ISA_REGION_LOG2": str(6 + c % 5)},
                       {"JP_BWT_FWD_CTXKEYS": "1", "JP_BWT_FWD_BYPASS": "0", "JP_BWT_FWD_PACKED": str(c & 1), "JP_BWT_FWD_PERIODIC": str((c >> 1) & 1), "JP_BWT_FWD_EMIT_REGION_LOG2": str(9 + c % 3)},
                       {"JP_BWT_FWD_CTXKEYS": "1", "JP_BWT_FWD_BYPASS": "0", "JP_BWT_FWD_KEYPASSES": str(4 + c % 5)}):
                for k in SWITCHES:
                    os.environ.pop(k, None)
                os.environ.update(fv)
                rc, got, _, _ = simt.forward(T)
                if rc != 0 or not (got == want).all():
                    fails += 1; print(f"case {c}: forward mismatch {fv} n={T.size} rc={rc}", flush=True)
        variants = [{}]
        if big:
            variants += [{"JP_BWT_INV_SINGLE": "1"}, {"JP_BWT_INV_SINGLE": "1", "JP_BWT_INV_LOG2M": "4"}, {"JP_BWT_INV_SINGLE": "1", "JP_BWT_INV_RANK_PLAN": ["i0", "s0", "s1,i2,s0"][c % 3]}]
        for v in variants:
            for k in SWITCHES:
                os.environ.pop(k, None)
            os.environ.update(v)
            rc, out, chunks, _ = simt.inverse(want, consume=bool(c & 1))
            if rc != 0 or not (out == T).all() or (("JP_BWT_INV_SINGLE" in v) and chunks == 0):
                fails += 1; print(f"case {c}: inverse mismatch {v} n={T.size} rc={rc} chunks={chunks}", flush=True)
    for k, val in saved.items():
        if val is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = val
    print(f"EMU FUZZ {cases} cases, {fails} failures, {time.time() - t0:.0f} s")
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
