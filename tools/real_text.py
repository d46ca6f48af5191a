"""Real-world input check: a corpus of source text found on the box (Python's standard library and site-packages),
concatenated to one block, through forward and inverse against the compiled reference. Real text has what the
synthetic generators lack: a heavy tail of long repeats (licence headers, generated tables, vendored copies).
    python tools/real_text.py [MiB]
Test/measurement infrastructure, not product."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import jampack_b200 as jp  # noqa: E402
import oracle  # noqa: E402


def corpus(limit):
    roots = ["/usr/lib/python3.12", "/usr/lib/python3", os.path.dirname(os.path.dirname(np.__file__))]
    out, size = [], 0
    for root in roots:
        for dp, _, files in os.walk(root):
            for f in sorted(files):
                if not f.endswith((".py", ".pyi", ".txt", ".h", ".c", ".json", ".rst", ".md", ".cfg")):
                    continue
                try:
                    b = open(os.path.join(dp, f), "rb").read()
                except OSError:
                    continue
                out.append(b); size += len(b)
                if size >= limit:
                    return np.frombuffer(b"".join(out)[:limit], dtype=np.uint8).copy()
    return np.frombuffer(b"".join(out), dtype=np.uint8).copy()


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    T = corpus(mib << 20)
    print(f"corpus {T.size / 2**20:.1f} MiB, {np.unique(T).size} distinct bytes")
    impl = "ref" if oracle.ref() is not None else "port"
    t0 = time.time(); want = oracle.forward(T, impl); t_ref = time.time() - t0
    for rep in range(3):
        got = jp.forward(T); fs = jp.last_stats().asdict()
    ok_f = bool((got == want).all())
    for rep in range(3):
        back = jp.inverse(want); s = jp.last_stats().asdict()
    ok_i = bool((back == T).all())
    print(f"forward ok={ok_f}: {fs['ms_total']:.2f} ms = {T.size / fs['ms_total'] / 1e6:.2f} GB/s (reference, 1 core+sssort threads: {t_ref:.1f} s); "
          f"rounds={fs['rounds']} depth={fs['initial_depth']} sum_a={sum(fs['active_fraction']):.3f} large={fs['large_fraction']:.3f} radix_tiles={fs['radix_tiles']} "
          f"bypass={fs['bypass_suffixes']}/{fs['bypass_runs']} period={fs['period']} ws={fs['device_bytes'] / T.size:.2f}N phases={[round(x, 3) for x in fs['ms_phase'][:5]]}")
    print(f"inverse ok={ok_i}: {s['ms_total']:.2f} ms = {T.size / s['ms_total'] / 1e6:.2f} GB/s phases={s['ms_phase'][:5]}")
    return 0 if ok_f and ok_i else 1


if __name__ == "__main__":
    sys.exit(main())
