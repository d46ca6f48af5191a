"""GPU bring-up ladder: runs every stage of the CUDA path against the oracle from tiny to 64 MiB and prints
enough detail per failure (first mismatch, counts) to debug from a log. Test infrastructure, not product.
    python tools/gpu_diag.py [--big] [--only fwd|inv]
"""
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import jampack_b200 as jp  # noqa: E402
import oracle  # noqa: E402

MiB = 1 << 20
ONLY = None
if "--only" in sys.argv:
    ONLY = sys.argv[sys.argv.index("--only") + 1]


def diff(name, got, want):
    got = np.asarray(got); want = np.asarray(want)
    if got.shape != want.shape:
        print(f"  [{name}] SHAPE {got.shape} vs {want.shape}"); return False
    bad = np.nonzero(got != want)[0]
    if bad.size == 0:
        print(f"  [{name}] ok ({got.size})"); return True
    i = int(bad[0])
    print(f"  [{name}] MISMATCH {bad.size}/{got.size}, first at {i}: got {got[i:i+8].tolist()} want {want[i:i+8].tolist()}; last at {int(bad[-1])}")
    return False


def step(title, fn):
    print(f"== {title}", flush=True)
    t = time.time()
    try:
        ok = fn()
    except Exception:
        traceback.print_exc(); ok = False
    print(f"   -> {'PASS' if ok else 'FAIL'} in {time.time()-t:.2f}s", flush=True)
    return ok


def ref_forward(T):
    return oracle.forward(T, "ref" if oracle.ref() is not None else "port")


def check_lf(kind, n, seed):
    def f():
        T = oracle.gen(kind, n, seed)
        B = ref_forward(T)
        nlen = n - n % 120
        idx = int(oracle.indices(B)[0])
        Map, Ct = oracle.build_map(B[:nlen], nlen, idx)
        lf, ct = jp.debug_lf(B[:nlen])
        ok = diff("ctable", ct, Ct)
        # Map[lf[i]-1] == i + (i >= idx)
        i = np.arange(nlen)
        ok &= diff("lf", Map[lf - 1], i + (i >= idx))
        return ok
    return f


def check_inv(kind, n, seed):
    def f():
        T = oracle.gen(kind, n, seed)
        B = ref_forward(T)
        out = jp.inverse(B)
        st = jp.last_stats().asdict()
        print("  stats", {k: st[k] for k in ("nlen", "kernel_launches", "subchains", "subchain_spacing", "ms_total", "ms_phase", "ms_h2d", "ms_d2h", "device_bytes")})
        return diff("inverse", out, T)
    return f


def check_sa(kind, n, seed):
    def f():
        T = oracle.gen(kind, n, seed)
        want = oracle.suffix_array(T)
        got = jp.debug_suffix_array(T)
        return diff("sa", got, want)
    return f


def check_fwd(kind, n, seed):
    def f():
        T = oracle.gen(kind, n, seed)
        want = ref_forward(T)
        got = jp.forward(T)
        st = jp.last_stats().asdict()
        print("  stats", {k: st[k] for k in ("nlen", "kernel_launches", "rounds", "symbol_bits", "initial_depth", "ms_total", "ms_phase", "ms_h2d", "ms_d2h", "device_bytes", "active_fraction")})
        nlen = n - n % 120
        ok = diff("bwt", got[:n], want[:n])
        if nlen:
            ok &= diff("indices", oracle.indices(got), oracle.indices(want))
        return ok
    return f


def gather():
    for tb, label in ((64 * MiB, "64MiB(L2)"), (256 * MiB, "256MiB"), (1024 * MiB, "1GiB")):
        for chains in (1 << 16, 1 << 18, 148 * 2048, 1 << 20):
            r = jp.debug_gather_rate(tb, chains, 256, True)
            print(f"  dependent  table={label:10s} chains={chains:8d}: {r/1e9:7.2f} G gathers/s = {r*32/1e9:8.1f} GB/s of sectors")
        r = jp.debug_gather_rate(tb, 1 << 20, 64, False)
        print(f"  independent table={label:10s} threads={1<<20}: {r/1e9:7.2f} G gathers/s = {r*32/1e9:8.1f} GB/s of sectors")
    return True


def main():
    print(jp.lib().jp_bwt_version().decode(), "devices:", jp.device_count(), "ref:", oracle.ref() is not None, flush=True)
    small = [("kat_quadratic", 240, 0), ("kat_quadratic", 250, 0), ("alla", 360, 0), ("kat_extremes", 240, 0),
             ("kat_quadratic", 119, 0), ("markov2", 4093, 9), ("repetitive", 70000, 3), ("uniform", 5000, 4)]
    medium = [("markov2", MiB, 1), ("uniform", MiB, 2), ("repetitive", MiB, 3), ("alla", MiB, 0), ("markov2", 8 * MiB, 1)]
    big = [("markov2", 64 * MiB, 1), ("uniform", 64 * MiB, 2), ("repetitive", 64 * MiB, 3), ("alla", 64 * MiB, 0)]
    res = []
    if "--sanitize" in sys.argv:      # small cases only: meant to run under compute-sanitizer
        medium = [("markov2", 300000, 1), ("alla", 200000, 0), ("repetitive", 400000, 3)]
        import numpy as np
        rng = np.random.default_rng(3)
        words = [bytes(rng.integers(97, 123, rng.integers(2, 9)).astype(np.uint8)) for _ in range(50)]
    if ONLY in (None, "inv"):
        for c in small + medium[:3]:
            if c[1] >= 120:
                res.append(step(f"LF table {c}", check_lf(*c)))
        for c in small + medium:
            res.append(step(f"inverse {c}", check_inv(*c)))
    if ONLY in (None, "fwd"):
        for c in small + medium[:4]:
            res.append(step(f"suffix array {c}", check_sa(*c)))
        for c in small + medium:
            res.append(step(f"forward {c}", check_fwd(*c)))
    if "--big" in sys.argv:
        for c in big:
            if ONLY in (None, "inv"):
                res.append(step(f"inverse {c}", check_inv(*c)))
                res.append(step(f"inverse again (warm) {c}", check_inv(*c)))
            if ONLY in (None, "fwd"):
                res.append(step(f"forward {c}", check_fwd(*c)))
    if "--gather" in sys.argv:
        res.append(step("gather micro-benchmark", gather))
    print("SUMMARY", sum(res), "/", len(res), "passed")
    return 0 if all(res) else 1


if __name__ == "__main__":
    sys.exit(main())
