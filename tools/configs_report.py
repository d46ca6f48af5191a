"""One table for BASELINE.json configs 2, 3 and 5 (stage level, one block at a time, device-resident) plus a block of
real source text: forward and inverse time, MB/s, rounds, sum of active fractions, workspace bytes per byte of block.
    python tools/configs_report.py > profiles/configs_r01.json
Measurement infrastructure, not product."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402
import jampack_b200 as jp  # noqa: E402
import synth  # noqa: E402
from real_text import corpus  # noqa: E402

MiB = 1 << 20
CASES = [("markov2", 64, 1), ("uniform", 64, 2), ("repetitive", 64, 3), ("alla", 64, 0), ("markov2", 256, 5), ("source-text", 64, 0)]
gold = {(c["kind"], c["len"], c["seed"]): c["fnv_all"] for c in json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "kat.json")))["big"]}
rows = []
for kind, mib, seed in CASES:
    T = corpus(mib * MiB) if kind == "source-text" else synth.gen(kind, mib * MiB, seed)
    n = T.size
    d_T = torch.from_numpy(T).cuda()
    d_B = torch.zeros(n + 480, dtype=torch.uint8, device="cuda")
    d_back = torch.zeros(n, dtype=torch.uint8, device="cuda")
    best_f = best_i = None
    for rep in range(4):
        jp.forward_device(d_T, d_B); f = jp.last_stats().asdict()
        jp.inverse_device(d_B, d_back); i = jp.last_stats().asdict()
        if rep and (best_f is None or f["ms_total"] < best_f["ms_total"]): best_f = f
        if rep and (best_i is None or i["ms_total"] < best_i["ms_total"]): best_i = i
    d_tmp = d_B.clone()
    d_back2 = jp.inverse_device(d_tmp, consume=True); ic = jp.last_stats().asdict()   # the 6N variant: the input block is scratch
    consume_ok = bool(torch.equal(d_back2[:n], d_T))
    del d_tmp, d_back2
    B = d_B.cpu().numpy()
    key = (kind, n, seed)
    rows.append({"input": f"{kind}({mib} MiB, seed {seed})", "bytes": n,
                 "forward_ms": round(best_f["ms_total"], 3), "forward_MBps": round(n / best_f["ms_total"] / 1e3, 1),
                 "rounds": best_f["rounds"], "initial_depth": best_f["initial_depth"],
                 "sum_active_fraction": round(sum(best_f["active_fraction"]), 4), "active_fraction": best_f["active_fraction"],
                 "forward_workspace_per_byte": round(best_f["device_bytes"] / n, 2), "forward_phases_ms": best_f["ms_phase"][:5],
                 "inverse_ms": round(best_i["ms_total"], 3), "inverse_MBps": round(n / best_i["ms_total"] / 1e3, 1),
                 "inverse_workspace_per_byte": round(best_i["device_bytes"] / n, 3), "inverse_phases_ms": best_i["ms_phase"][:5],
                 "inverse_workspace_per_byte_consumed_input": round(ic["device_bytes"] / n, 3),
                 "inverse_stream_bytes_per_byte": round(abs(best_i["stream_chunks"]) * 1024 / n, 3),
                 "round_trip": bool(torch.equal(d_back, d_T)) and consume_ok,
                 "forward_matches_reference_hash": (("%016x" % synth.fnv(B)) == gold[key]) if key in gold else None})
    del d_T, d_B, d_back
    torch.cuda.empty_cache()
print(json.dumps({"note": "one block at a time on one stream, blocks resident in HBM; best of 3 warm runs; 1 x B200", "rows": rows}, indent=1))
