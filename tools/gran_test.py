import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, jampack_b200 as jp, synth
MiB = 1 << 20
print("L2_FETCH", os.environ.get("JP_BWT_L2_FETCH"))
for tb in (256 * MiB, 1024 * MiB):
    for dep in (True, False):
        r = jp.debug_gather_rate(tb, 148 * 2048, 256, dep)
        print(f"  table {tb>>20} MiB dependent={dep}: {r/1e9:.2f} G/s = {r*32/1e9:.0f} GB/s sectors")
n = 64 * MiB
T = synth.gen("markov2", n, 1)
d_T = torch.from_numpy(T).cuda(); d_B = torch.zeros(n + 480, dtype=torch.uint8, device="cuda"); d_back = torch.zeros(n, dtype=torch.uint8, device="cuda")
for i in range(3):
    jp.forward_device(d_T, d_B); f = jp.last_stats().asdict()
    jp.inverse_device(d_B, d_back); s = jp.last_stats().asdict()
print("  inverse", s["ms_total"], s["ms_phase"][:5]); print("  forward", f["ms_total"], f["ms_phase"][:5])
assert torch.equal(d_back, d_T)
