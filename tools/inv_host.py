import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, jampack_b200 as jp, synth
n = 64 << 20
T = synth.gen("markov2", n, 1)
B = jp.forward(T)
for i in range(3):
    out = jp.inverse(B); s = jp.last_stats().asdict()
print("host api (consume):", s["ms_total"], s["ms_phase"][:5], s["device_bytes"])
d = torch.from_numpy(B).cuda()
for i in range(3):
    jp.inverse_device(d.clone(), consume=True); s2 = jp.last_stats().asdict()
print("device api consume:", s2["ms_total"], s2["ms_phase"][:5], s2["device_bytes"])
for i in range(3):
    jp.inverse_device(d); s3 = jp.last_stats().asdict()
print("device api const  :", s3["ms_total"], s3["ms_phase"][:5], s3["device_bytes"])
