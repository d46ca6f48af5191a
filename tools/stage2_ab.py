"""Timing of the second stage's first half on a block resident in HBM: forward_device -> src_rle0_device.
    python tools/stage2_ab.py [kind] [MiB]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, jampack_b200 as jp, synth
kind = sys.argv[1] if len(sys.argv) > 1 else "markov2"
mib = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n = mib << 20
d_T = torch.from_numpy(synth.gen(kind, n, 1)).cuda()
d_B = jp.forward_device(d_T)
best = None
for i in range(5):
    freq, rle, rlen = jp.src_rle0_device(d_B); s = jp.last_stats().asdict()
    if i >= 1 and (best is None or s["ms_total"] < best["ms_total"]): best = s
print(f"{kind} {mib}MiB src+rle0 total={best['ms_total']:.3f} ms phases(tables, ranks, rle0)={[round(x,3) for x in best['ms_phase'][:3]]} symbols={int(rlen.sum())} ws={best['device_bytes']/n:.2f}N -> {n/best['ms_total']/1e6:.2f} GB/s")
