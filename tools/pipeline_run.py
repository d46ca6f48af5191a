"""BASELINE.json configs[3]: a 1 GiB corpus of 16 x 64 MiB markov2 blocks through the reference's own CLI
(-b64 -t16), once unmodified (CPU BWT stage) and once with our stage linked in place of bwt.cpp
(oracle/_ref/Jampack_shim), with the stage's wall-clock trace. The .jam files must be byte-identical.
    python tools/pipeline_run.py [--blocks 16] [--devices 0,1] [--skip-ref]
Test/measurement infrastructure, not product."""
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import synth  # noqa: E402

MiB = 1 << 20


def arg(name, default):
    return sys.argv[sys.argv.index(name) + 1] if name in sys.argv else default


def run(cmd, env=None):
    t0 = time.perf_counter()
    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-2000:])
        raise SystemExit(f"{cmd[0]} failed rc={r.returncode}")
    return dt, r.stderr


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        while True:
            b = f.read(1 << 24)
            if not b:
                break
            h.update(b)
    return h.hexdigest()


def stage_summary(calls):
    """Per direction, from the library's per-call trace lines: bytes / UNION of the calls' wall-clock intervals (SURVEY.md 8d,
    config 4: "sum of block bytes / union of stage wall time"), the longest wait for a context and the call times."""
    out = {}
    for d, name in ((0, "forward"), (1, "inverse")):
        iv = []
        for ln in calls:
            kv = dict(x.split("=") for x in ln.split())
            if int(kv["dir"]) != d:
                continue
            iv.append((int(kv["t0_us"]), int(kv["t0_us"]) + int(kv["call_us"]), int(kv["len"]), int(kv["wait_us"]), int(kv["call_us"]), int(kv["dev"]),
                       float(kv["h2d_ms"]), float(kv["kernels_ms"]), float(kv["d2h_ms"])))
        if not iv:
            continue
        iv.sort()
        union, cur_a, cur_b = 0, iv[0][0], iv[0][1]
        for a, b, *_ in iv[1:]:
            if a > cur_b:
                union += cur_b - cur_a; cur_a, cur_b = a, b
            else:
                cur_b = max(cur_b, b)
        union += cur_b - cur_a
        nbytes = sum(x[2] for x in iv)
        call = sorted(x[4] for x in iv)
        out[name] = {"calls": len(iv), "MB": round(nbytes / 1e6, 1), "union_s": round(union / 1e6, 4), "MBps_over_union": round(nbytes / union, 1),
                     "max_wait_ms": round(max(x[3] for x in iv) / 1e3, 2), "call_ms_median": round(call[len(call) // 2] / 1e3, 2), "call_ms_max": round(call[-1] / 1e3, 2),
                     "h2d_ms_median": sorted(x[6] for x in iv)[len(iv) // 2], "kernels_ms_median": sorted(x[7] for x in iv)[len(iv) // 2],
                     "d2h_ms_median": sorted(x[8] for x in iv)[len(iv) // 2], "devices_used": sorted(set(x[5] for x in iv))}
    return out


def main():
    blocks = int(arg("--blocks", "16"))
    threads = arg("--threads", "16")
    devices = arg("--devices", None)
    ref = os.path.join(ROOT, "oracle", "_ref", "Jampack_ref")
    shim = os.path.join(ROOT, "oracle", "_ref", "Jampack_shim")
    d = tempfile.mkdtemp(prefix="jp_pipeline_")
    src = os.path.join(d, "corpus.bin")
    with open(src, "wb") as f:
        for b in range(blocks):
            synth.gen("markov2", 64 * MiB, 100 + b).tofile(f)
    out = {"blocks": blocks, "block_mib": 64, "flags": f"-b64 -t{threads}", "host_cores": os.cpu_count(), "devices": devices or "all",
           "effective_threads": min(int(threads), os.cpu_count() or 1)}      # Opt.Threads is clamped to the core count (jampack.cpp:188-189)
    env = dict(os.environ, JP_BWT_TRACE="2")          # per-call lines: the union-of-stage-time metric needs them
    keep_calls = "--calls" in sys.argv
    if devices:
        env["JP_BWT_DEVICES"] = devices
    jam_s, back_s = os.path.join(d, "shim.jam"), os.path.join(d, "shim.back")
    dt, err = run([shim, "c", src, jam_s, "-b64", f"-t{threads}"], env)
    out["shim_compress_s"] = round(dt, 2)
    out["shim_compress_trace"] = re.findall(r"\[jp_bwt trace\] (.*)", err)
    calls = re.findall(r"\[jp_bwt call\] (.*)", err)
    out["compress_stage"] = stage_summary(calls)
    if calls and keep_calls:
        out["compress_calls"] = calls
    dt, err = run([shim, "d", jam_s, back_s, f"-t{threads}"], env)
    out["shim_decompress_s"] = round(dt, 2)
    out["shim_decompress_trace"] = re.findall(r"\[jp_bwt trace\] (.*)", err)
    calls = re.findall(r"\[jp_bwt call\] (.*)", err)
    out["decompress_stage"] = stage_summary(calls)
    if calls and keep_calls:
        out["decompress_calls"] = calls
    out["shim_round_trip"] = sha(back_s) == sha(src)
    out["jam_bytes"] = os.path.getsize(jam_s)
    out["shim_jam_sha256"] = sha(jam_s)
    if "--skip-ref" not in sys.argv:
        jam_r, back_r = os.path.join(d, "ref.jam"), os.path.join(d, "ref.back")
        dt, _ = run([ref, "c", src, jam_r, "-b64", f"-t{threads}"])
        out["ref_compress_s"] = round(dt, 2)
        dt, _ = run([ref, "d", jam_r, back_r, f"-t{threads}"])
        out["ref_decompress_s"] = round(dt, 2)
        out["ref_jam_sha256"] = sha(jam_r)
        out["jam_identical"] = out["ref_jam_sha256"] == out["shim_jam_sha256"]
        # cross decode: each binary reads the other's stream
        dt, _ = run([shim, "d", jam_r, back_s, f"-t{threads}"], env)
        out["shim_decodes_ref"] = sha(back_s) == sha(src)
    print(json.dumps(out, indent=1))
    for f in os.listdir(d):
        os.remove(os.path.join(d, f))
    os.rmdir(d)


if __name__ == "__main__":
    main()
