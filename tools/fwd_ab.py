import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, jampack_b200 as jp, synth, json
MiB = 1 << 20
kind = sys.argv[1] if len(sys.argv) > 1 else "markov2"
mib = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n = mib * MiB
T = synth.gen(kind, n, {"markov2": 5 if mib == 256 else 1, "uniform": 2, "repetitive": 3}.get(kind, 0))
d_T = torch.from_numpy(T).cuda(); d_B = torch.zeros(n + 480, dtype=torch.uint8, device="cuda")
best = None
for i in range(5):
    jp.forward_device(d_T, d_B); s = jp.last_stats().asdict()
    if i >= 2 and (best is None or s["ms_total"] < best["ms_total"]): best = s
fnv = "%016x" % synth.fnv(d_B.cpu().numpy())
gold = {c["name"]: c["fnv_all"] for c in json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "kat.json")))["big"]}
ok = gold.get(f"{kind}-{mib}M") == fnv
print(f"ENV={ {k:v for k,v in os.environ.items() if k.startswith('JP_BWT')} } {kind} {mib}MiB golden_ok={ok} total={best['ms_total']:.3f} phases={[round(x,3) for x in best['ms_phase'][:5]]} rounds={best['rounds']} large={best['large_fraction']:.3f} radix_tiles={best['radix_tiles']} bypass={best['bypass_suffixes']}/{best['bypass_runs']} period={best['period']} a={best['active_fraction']} ws={best['device_bytes']/n:.2f}N -> {n/best['ms_total']/1e6:.2f} GB/s")
