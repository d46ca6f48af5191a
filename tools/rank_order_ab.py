"""Ranking kernel of the single-walk inverse: index order against scattered block order on inputs with a growing share of
single-symbol runs (the probe of k_inv_rank_probe decides between them). Measurement infrastructure, not product."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, jampack_b200 as jp, synth
MiB = 1 << 20
n = 64 * MiB
rng = np.random.default_rng(5)
def mix(frac, pieces):
    T = synth.gen("markov2", n, 1).copy()
    L = int(n * frac / pieces)
    for k in range(pieces):
        a = int((k + 0.5) * n / pieces) - L // 2
        T[a:a + L] = 0
    return T
cases = [("markov2", synth.gen("markov2", n, 1)), ("alla", synth.gen("alla", n, 0)), ("repetitive", synth.gen("repetitive", n, 3)), ("uniform", synth.gen("uniform", n, 2))]
for frac, pieces in ((0.002, 1), (0.01, 1), (0.01, 64), (0.05, 4), (0.2, 2), (0.5, 1)):
    cases.append((f"markov2 + {frac:.3f} zeros in {pieces}", mix(frac, pieces)))
try:
    from real_text import corpus
    cases.append(("source text", corpus(n)))
except Exception as e:
    print("no corpus", e)
for name, T in cases:
    d_T = torch.from_numpy(np.ascontiguousarray(T)).cuda(); d_B = torch.zeros(T.size + 480, dtype=torch.uint8, device="cuda"); d_back = torch.zeros(T.size, dtype=torch.uint8, device="cuda")
    jp.forward_device(d_T, d_B)
    out = []
    for plan in (sys.argv[1:] or ["i0", "s0", "s2,i0", "s4,i0", "i2,s0", "i4,s0", "s4,i8,s0"]):
        os.environ["JP_BWT_INV_RANK_PLAN"] = plan
        best = 1e9
        for i in range(4):
            jp.inverse_device(d_B, d_back); s = jp.last_stats().asdict(); best = min(best, s["ms_phase"][3])
        out.append(f"{plan} {best:.3f}")
    ok = torch.equal(d_back, d_T)
    print(f"{name:34s} ok={ok} rank ms: " + "  ".join(out), flush=True)
