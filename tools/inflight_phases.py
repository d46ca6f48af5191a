"""Per-phase times of the inverse with K blocks in flight (K caller threads, one stream and one block each): which
phases stretch when calls overlap. Measurement infrastructure, not product.
    python tools/inflight_phases.py [K ...]"""
import os, sys, threading
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, jampack_b200 as jp, synth
MiB = 1 << 20
n = 64 * MiB
T = synth.gen("markov2", n, 1)
d_T = torch.from_numpy(T).cuda(); d_B = torch.zeros(n + 480, dtype=torch.uint8, device="cuda")
jp.forward_device(d_T, d_B)
names = ["hist", "lf", "walk", "rank", "place"]
for K in [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4]:
    ins = [d_B.clone() for _ in range(K)]; outs = [torch.zeros(n, dtype=torch.uint8, device="cuda") for _ in range(K)]
    streams = [torch.cuda.Stream() for _ in range(K)]
    res = [[] for _ in range(K)]
    reps = 12
    bar = threading.Barrier(K)
    def work(k):
        with torch.cuda.stream(streams[k]):
            for r in range(reps):
                if r == 4: bar.wait()
                jp.inverse_device(ins[k], outs[k])
                s = jp.last_stats().asdict()
                if r >= 4: res[k].append([s["ms_total"]] + list(s["ms_phase"][:5]))
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    th = [threading.Thread(target=work, args=(k,)) for k in range(K)]
    t0 = time.time()
    for t in th: t.start()
    for t in th: t.join()
    torch.cuda.synchronize()
    a = np.array([x for r in res for x in r]).mean(axis=0)
    ok = all(torch.equal(o, d_T) for o in outs)
    print(f"K={K} ok={ok} per call: total {a[0]:.3f} ms  " + "  ".join(f"{nm} {v:.3f}" for nm, v in zip(names, a[1:])) + f"   -> {a[0] / K:.3f} ms per block if fully overlapped")
