"""Single-walk inverse against the two-pass inverse: parity on a ladder of inputs, then timing on 64 MiB.
    python tools/inv_single_ab.py [--quick]
Test infrastructure (drives the product through its C-ABI; the forward transform supplies the inputs)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, jampack_b200 as jp, synth
MiB = 1 << 20


def run(d_B, n, mode, consume, reps):
    if mode is None: os.environ.pop("JP_BWT_INV_SINGLE", None)
    else: os.environ["JP_BWT_INV_SINGLE"] = str(mode)
    d_back = torch.zeros(n, dtype=torch.uint8, device="cuda")
    best = None
    for i in range(reps):
        src = d_B.clone() if consume else d_B
        d_back.zero_()
        jp.inverse_device(src, d_back, consume=consume); s = jp.last_stats().asdict()
        if best is None or s["ms_total"] < best["ms_total"]: best = s
    return d_back, best


fails = 0
cases = [("markov2", 65536, 1), ("uniform", 120 * 999, 2), ("markov2", 1 * MiB + 77, 3), ("repetitive", 3 * MiB, 4), ("alla", 2 * MiB + 5, 5),
         ("markov2", 8 * MiB, 6), ("uniform", 16 * MiB + 1234, 7), ("kat_quadratic", 1 * MiB, 8), ("markov2", 64 * MiB, 1)]
if "--quick" not in sys.argv:
    cases += [("uniform", 64 * MiB, 2), ("repetitive", 64 * MiB, 3), ("alla", 48 * MiB + 7, 4), ("markov2", 256 * MiB, 5)]
for kind, n, seed in cases:
    try:
        T = synth.gen(kind, n, seed)
    except Exception as e:
        print("skip", kind, e); continue
    d_T = torch.from_numpy(T).cuda(); d_B = torch.zeros(n + 480, dtype=torch.uint8, device="cuda")
    jp.forward_device(d_T, d_B)
    for mode, consume in ((0, False), (1, False), (1, True), (None, True)):
        back, st = run(d_B, n, mode, consume, 3)
        ok = bool(torch.equal(back, d_T))
        fails += 0 if ok else 1
        nb = int((back != d_T).sum().item()) if not ok else 0
        print(f"{kind:10s} n={n:10d} single={mode} consume={int(consume)} ok={ok} bad={nb} chunks={st['stream_chunks']} ({abs(st['stream_chunks'])*1024/max(n,1):.3f} n) "
              f"total={st['ms_total']:.3f} phases={[round(x,3) for x in st['ms_phase'][:5]]} bytes={st['device_bytes']} -> {n/st['ms_total']/1e6:.2f} GB/s", flush=True)
    # host entry point (pinned), as the reference's caller would drive it
    os.environ.pop("JP_BWT_INV_SINGLE", None)
    if n <= 64 * MiB:
        hb = jp.inverse(d_B.cpu().numpy())
        ok = bool(np.array_equal(hb, T)); fails += 0 if ok else 1
        print(f"{kind:10s} n={n:10d} host entry ok={ok} chunks={jp.last_stats().asdict()['stream_chunks']}", flush=True)
    del d_T, d_B
# walker-count sweep on the headline block
T = synth.gen("markov2", 64 * MiB, 1); n = T.size
d_T = torch.from_numpy(T).cuda(); d_B = torch.zeros(n + 480, dtype=torch.uint8, device="cuda"); jp.forward_device(d_T, d_B)
for per_sm in (8, 6, 5, 4, 3):
    os.environ["JP_BWT_INV_WBLOCKS_PER_SM"] = str(per_sm)
    back, st = run(d_B, n, None, True, 5)
    print(f"walker blocks/SM={per_sm} ok={bool(torch.equal(back, d_T))} chunks={st['stream_chunks']} ({abs(st['stream_chunks'])*1024/n:.3f} n) total={st['ms_total']:.3f} phases={[round(x,3) for x in st['ms_phase'][:5]]}", flush=True)
os.environ.pop("JP_BWT_INV_WBLOCKS_PER_SM", None)
for l2m in (3, 5, 6, 4):
    os.environ["JP_BWT_INV_LOG2M"] = str(l2m)
    back, st = run(d_B, n, None, True, 5)
    print(f"log2m={l2m} ok={bool(torch.equal(back, d_T))} chunks={st['stream_chunks']} ({abs(st['stream_chunks'])*1024/n:.3f} n) total={st['ms_total']:.3f} phases={[round(x,3) for x in st['ms_phase'][:5]]}", flush=True)
os.environ.pop("JP_BWT_INV_LOG2M", None)
print("FAILS", fails)
sys.exit(1 if fails else 0)
