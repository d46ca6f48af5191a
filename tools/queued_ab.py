"""Runs the A/B experiments that were built but not yet measured (they are all off by default):
  * JP_BWT_FWD_RUNSKIP=1   forward, single-symbol runs ordered by run length (bwt_forward.cu "run skip")
  * JP_BWT_INV_LF_BLOCKS=4 inverse LF build capped at 64 registers (4 blocks per SM)
  * JP_BWT_INV_RANK_ILP=4  sub-chain ranking with four nodes per thread
  * JP_BWT_INV_ILP=4       single-walk inverse with four sub-chains per walker thread (a quarter of the warps)
Each case runs in its own process (the switches are read per call, but a fresh process keeps the arenas comparable).
    python tools/queued_ab.py            # on a B200 box: prints one line per case, parity checked against golden hashes"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env, script, *args):
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", script), *args], cwd=ROOT, env=e, capture_output=True, text=True, timeout=900)
    out = [ln for ln in r.stdout.splitlines() if ln.strip()]
    print((out[-1] if out else "(no output) " + r.stderr[-300:]), flush=True)


if __name__ == "__main__":
    print("== parity of the experimental paths on the GPU")
    e = dict(os.environ, JP_BWT_TEST_EXPERIMENTAL="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "-q", "-m", "gpu", "-k", "run_skip or four_chains"], cwd=ROOT, env=e,
                       capture_output=True, text=True, timeout=1800)
    print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:])
    print("== forward, run skip off / on")
    for kind, mib in (("markov2", "64"), ("uniform", "64"), ("alla", "64"), ("repetitive", "64")):
        for v in ("0", "1"):
            run({"JP_BWT_FWD_RUNSKIP": v}, "fwd_ab.py", kind, mib)
    print("== inverse, LF build 3 / 4 blocks per SM")
    for v in ("3", "4"):
        run({"JP_BWT_INV_LF_BLOCKS": v}, "inv_ab.py", "markov2", "64")
    print("== inverse, ranking 1 / 4 nodes per thread")
    for v in ("1", "4"):
        run({"JP_BWT_INV_RANK_ILP": v}, "inv_ab.py", "markov2", "64")
    print("== inverse one block at a time: one chain per lane (5 and 2 blocks of 8 warps per SM) / four chains per lane (1, 2, 3 blocks of 4 warps per SM)")
    for env in ({}, {"JP_BWT_INV_WBLOCKS_PER_SM": "2"}, {"JP_BWT_INV_ILP": "4", "JP_BWT_INV_WBLOCKS_PER_SM": "1"}, {"JP_BWT_INV_ILP": "4"},
                {"JP_BWT_INV_ILP": "4", "JP_BWT_INV_WBLOCKS_PER_SM": "3"}):
        print(env); run(env, "inv_ab.py", "markov2", "64")
    print("== inverse, four blocks in flight (bench.py value / e2e): the same five configurations")
    for env in ({}, {"JP_BWT_INV_WBLOCKS_PER_SM": "2"}, {"JP_BWT_INV_ILP": "4", "JP_BWT_INV_WBLOCKS_PER_SM": "1"}, {"JP_BWT_INV_ILP": "4"},
                {"JP_BWT_INV_ILP": "4", "JP_BWT_INV_WBLOCKS_PER_SM": "3"}):
        e = dict(os.environ); e.update(env)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "10", "--warmup", "3", "--no-cpu-baseline", "--no-forward"],
                           cwd=ROOT, env=e, capture_output=True, text=True, timeout=900)
        try:
            import json
            d = json.loads(r.stdout.strip().splitlines()[-1])
            print(env, "value", d["value"], "e2e", d["e2e"]["value"], "single", d["single_stream"]["value"], d["parity"])
        except Exception:  # noqa: BLE001
            print(env, "bench failed:", r.stderr[-300:])
