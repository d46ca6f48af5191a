"""Reads an `ncu --page raw --csv` dump and prints / stores per-kernel DRAM traffic (dram__bytes_read.sum +
dram__bytes_write.sum per launch) and durations.
    python tools/ncu_traffic.py raw.csv [--update profiles/traffic.json --key inverse_walk_single --sum k_inv_walk_stream,k_inv_rank_packed,...]
"""
import csv, json, re, sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
        "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}


def load(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h, u = rows[hi], rows[hi + 1]
    col = {name: h.index(name) for name in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}
    out = []
    for r in rows[hi + 2:]:
        if len(r) != len(h):
            continue
        def val(name):
            i = col[name]
            return float(r[i].replace(",", "")) * UNIT[u[i]]
        out.append({"kernel": re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("jp::", ""),
                    "read": val("dram__bytes_read.sum"), "write": val("dram__bytes_write.sum"), "ms": val("gpu__time_duration.sum")})
    return out


if __name__ == "__main__":
    rows = load(sys.argv[1])
    for r in rows:
        print(f"{r['kernel']:24s} {r['ms']:8.4f} ms  read {r['read']/1e6:10.1f} MB  write {r['write']/1e6:10.1f} MB")
    if "--sum-all" in sys.argv:                       # every launch of the dump (one whole call captured with cudaProfilerStart/Stop)
        path = sys.argv[sys.argv.index("--update") + 1]
        key = sys.argv[sys.argv.index("--key") + 1]
        d = json.load(open(path))
        d[key] = int(round(sum(r["read"] + r["write"] for r in rows), -5))
        d[key + "_launches"] = len(rows)
        d[key + "_ms_under_ncu"] = round(sum(r["ms"] for r in rows), 3)
        json.dump(d, open(path, "w"), indent=1)
        print("updated", path, key, d[key], "over", len(rows), "launches")
    elif "--update" in sys.argv:
        path = sys.argv[sys.argv.index("--update") + 1]
        key = sys.argv[sys.argv.index("--key") + 1]
        names = sys.argv[sys.argv.index("--sum") + 1].split(",")
        d = json.load(open(path))
        tot = 0
        for r in rows:
            if r["kernel"] in names:
                d[r["kernel"]] = int(round(r["read"] + r["write"], -5))
                tot += r["read"] + r["write"]
        d[key] = int(round(tot, -5))
        json.dump(d, open(path, "w"), indent=1)
        print("updated", path, key, d[key])
