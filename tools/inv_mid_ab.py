"""Two-pass against single-walk inverse on mid-size blocks (auto mode), consume entry point."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, jampack_b200 as jp, synth
MiB = 1 << 20
for kind, mib, seed in (("markov2", 24, 1), ("markov2", 32, 1), ("uniform", 32, 2), ("repetitive", 36, 3), ("markov2", 40, 4), ("markov2", 47, 5), ("markov2", 48, 6), ("alla", 33, 0)):
    n = mib * MiB + 77
    T = synth.gen(kind, n, seed)
    d_T = torch.from_numpy(T).cuda(); d_B = torch.zeros(n + 480, dtype=torch.uint8, device="cuda")
    jp.forward_device(d_T, d_B)
    for mode in ("0", None):
        if mode is None: os.environ.pop("JP_BWT_INV_SINGLE", None)
        else: os.environ["JP_BWT_INV_SINGLE"] = mode
        best = None
        for rep in range(4):
            src = d_B.clone()
            out = jp.inverse_device(src, consume=True); st = jp.last_stats().asdict()
            if rep and (best is None or st["ms_total"] < best["ms_total"]): best = st
        ok = bool(torch.equal(out[:n], d_T))
        print(f"{kind:10s} {mib:3d} MiB single={mode} ok={ok} chunks={best['stream_chunks']} ({abs(best['stream_chunks'])*1024/n:.3f} n) total={best['ms_total']:.3f} "
              f"phases={[round(x,3) for x in best['ms_phase'][:5]]} ws={best['device_bytes']/n:.3f}", flush=True)
    del d_T, d_B
