import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, jampack_b200 as jp, synth
MiB = 1 << 20
kind = sys.argv[1] if len(sys.argv) > 1 else "markov2"
mib = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n = mib * MiB
T = synth.gen(kind, n, 1)
d_T = torch.from_numpy(T).cuda(); d_B = torch.zeros(n + 480, dtype=torch.uint8, device="cuda"); d_back = torch.zeros(n, dtype=torch.uint8, device="cuda")
jp.forward_device(d_T, d_B)
best = None
for i in range(6):
    d_back.zero_()
    jp.inverse_device(d_B, d_back); s = jp.last_stats().asdict()
    if i >= 2 and (best is None or s["ms_total"] < best["ms_total"]): best = s
ok = torch.equal(d_back, d_T)
print(f"FLAGS={os.environ.get('JP_BWT_INV_FLAGS')} LOG2M={os.environ.get('JP_BWT_INV_LOG2M')} {kind} {mib}MiB ok={ok} total={best['ms_total']:.3f} phases={[round(x,3) for x in best['ms_phase'][:5]]} -> {n/best['ms_total']/1e6:.2f} GB/s")
