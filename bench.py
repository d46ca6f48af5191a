#!/usr/bin/env python
"""bench.py -- the BWT stage of Jampack on B200: MB/s of inverse (headline) and forward BWT on 64 MiB blocks.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): inverse BWT of one 64 MiB markov2(seed=1) block with all 120 stored
indices; one block per GPU per step (weak scaling: blocks are independent, no collective on the data path).
The forward transform of the same block (configs[2]/[3] shape) is reported in the "forward" object.

  value     MB/s (1e6 bytes of block per second) with the block resident in HBM, CUDA events around K steps
  e2e       the same through the host C-ABI (jp_bwt_inverse): pinned host block in, pinned host block out,
            both copies inside the timed region
  roofline  the inverse walk (decode walk + ranking + placement; on blocks under 30 Mi the two LF-walk kernels +
            ranking) against the measured HBM copy bandwidth of MEASURED_PEAKS.json, at SURVEY.md 8d's 64 B of
            random sectors per byte -- the figure is fixed by the problem, the single-walk path itself gathers
            one sector per byte; `rand_peak` is the random
            32 B-sector gather rate measured live by the library's own micro-benchmark
  cpu_baseline  the unmodified reference (oracle/_ref) on this box's host cores, block-parallel like
            Jampack::Compress/Decompress (jampack.cpp:215-219, :313-317): one block per core; `single_block` is the
            other all-core shape (one block, Opt.Threads = cores); `value` is the better of the two
  legacy_cuda_baseline  the reference's own CUDA inverse (bwt.cpp:8-19, :186-240), unmodified, in a child process

--impl reference times only the reference CPU implementation (rank 0; other ranks exit).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
MiB = 1 << 20
TRAILER = 480


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--block-mib", type=int, default=64)
    ap.add_argument("--kind", default="markov2")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-forward", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE.json configs (forward on uniform / repetitive / all-a, 256 MiB block, 8 MiB CPU round trip)")
    ap.add_argument("--callers", type=int, default=4, help="blocks in flight per GPU (the library keeps up to JP_BWT_MAX_CTX=4 contexts per device)")
    return ap.parse_args()


# ---- clocks -------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        busy = [x for x in sm if x > 0]
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(key):
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(p):
        try:
            return json.load(open(p)).get(key)
        except Exception:  # noqa: BLE001
            return None
    return None


# ---- the reference on the host cores (cpu_baseline leg and --impl reference) ---------------------------------
def cpu_reference(direction, block, fwd_out, cores, steps=1, warmup=0, budget_s=300.0):
    """Block-parallel shape: `cores` independent copies of the block, one reference call per core.
    Returns dict(value MB/s, ms_per_step, steps, kind, cores, sample)."""
    import oracle                                    # checker / baseline only -- never on the product path
    ref = oracle.ref()
    kind = "reference" if ref is not None else "port"
    n = block.size
    P = cores
    outs = [np.empty(n + TRAILER, dtype=np.uint8) for _ in range(P)]
    src = block if direction == "forward" else fwd_out
    pp = (C.c_void_p * P)(*[src.ctypes.data] * P)
    po = (C.c_void_p * P)(*[o.ctypes.data for o in outs])
    lens = (C.c_int32 * P)(*[src.size] * P)

    def one_step():
        if ref is not None:
            if direction == "forward":
                return ref.ref_bwt_forward_batch(pp, lens, po, P, P)
            return ref.ref_bwt_inverse_batch(pp, lens, po, P, P, 1)
        t0 = time.perf_counter()                     # the C restatement, serial per block, thread per block
        th = [threading.Thread(target=(oracle.forward if direction == "forward" else oracle.inverse), args=(src, "port"))
              for _ in range(P)]
        [t.start() for t in th]; [t.join() for t in th]
        return time.perf_counter() - t0

    times = []
    t_first = one_step() if warmup > 0 else None
    if t_first is not None and (warmup + steps) * t_first > budget_s:
        warmup, steps = 1, max(1, int(budget_s / t_first) - 1)
    for _ in range(max(warmup - 1, 0)):
        one_step()
    for _ in range(steps):
        times.append(one_step())
    want = block if direction == "inverse" else fwd_out
    ok = all((o[: want.size] == want).all() for o in outs[:2])
    sec = sum(times) / len(times)
    return {"value": round(P * n / sec / 1e6, 2), "unit": "MB/s", "cores": P, "kind": kind, "ms_per_step": round(sec * 1e3, 1),
            "steps": len(times), "output_matches": bool(ok),
            "sample": f"{P} x {n >> 20} MiB {direction} blocks, one per core (block-parallel, jampack.cpp:215-219), {len(times)} batch(es)"}


def cpu_reference_single_block(direction, block, fwd_out, cores):
    """Single-block shape (BASELINE.md plan item 4): ONE block with Opt.Threads = all cores. The inverse rounds that to
    its unit table (bwt.cpp:116-132: >= 16 threads -> 120 units on 30 threads); the forward parallelises only sssort
    (divsufsort.cpp:1491-1521) over the OpenMP default team."""
    import oracle
    ref = oracle.ref()
    if ref is None:
        return None
    t0 = time.perf_counter()
    if direction == "forward":
        out = oracle.forward(block, "ref")
        ok = bool((out == fwd_out).all())
    else:
        out = oracle.inverse(fwd_out, "ref", threads=cores)
        ok = bool((out == block).all())
    sec = time.perf_counter() - t0
    return {"value": round(block.size / sec / 1e6, 2), "unit": "MB/s", "threads": cores, "output_matches": ok,
            "sample": f"1 x {block.size >> 20} MiB {direction} block, Opt.Threads = {cores}"}


def legacy_cuda_baseline(T, B):
    """The reference's own CUDA inverse, in a child process with a timeout (see oracle/legacy_cuda.py)."""
    import tempfile
    from oracle import legacy_cuda
    if not legacy_cuda.available():
        return {"unavailable": "oracle/_ref/libjamref_cuda.so was not built (make -C oracle ref_cuda needs /root/reference and nvcc)"}
    with tempfile.TemporaryDirectory() as d:
        pb, pt = os.path.join(d, "B.npy"), os.path.join(d, "T.npy")
        np.save(pb, B); np.save(pt, T)
        try:
            r = subprocess.run([sys.executable, "-m", "oracle.legacy_cuda", pb, pt, "2"], cwd=ROOT, capture_output=True, text=True, timeout=90)
        except subprocess.TimeoutExpired:
            return {"unavailable": "legacy CUDA inverse did not finish within 90 s"}
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    if r.returncode != 0 or not lines:
        return {"unavailable": f"legacy CUDA inverse failed (exit {r.returncode}): {(r.stderr or r.stdout).strip()[-200:]}"}
    try:
        return json.loads(lines[-1])
    except ValueError:
        return {"unavailable": "legacy CUDA inverse printed no result"}


def run_reference(args, rank):
    if rank != 0:
        return 0
    import oracle
    import synth
    n = args.block_mib * MiB
    T = synth.gen(args.kind, n, args.seed)
    fwd = oracle.forward(T, "ref" if oracle.ref() is not None else "port")
    cores = os.cpu_count() or 1
    inv = cpu_reference("inverse", T, fwd, cores, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": "inv BWT MB/s", "value": inv["value"], "unit": "MB/s", "n_gpus": args.gpus,
            "steps": inv["steps"], "steps_requested": args.steps, "warmup": args.warmup, "ms_per_step": inv["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, n), "cpu_baseline": {k: inv[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": inv["value"], "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "host_cores": cores}
    if not args.no_forward:
        f = cpu_reference("forward", T, fwd, cores, steps=1, warmup=0)
        line["forward"] = {"value": f["value"], "unit": "MB/s", "ms_per_step": f["ms_per_step"],
                           "cpu_baseline": {k: f[k] for k in ("value", "unit", "cores", "kind", "sample")}}
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, n):
    return {"workload": f"inverse BWT of {args.block_mib} MiB {args.kind}(seed={args.seed}) blocks, all 120 stored primary indices each "
                        "(BASELINE.json configs[1]); 4 blocks in flight per GPU per step, as the reference's block loop keeps several "
                        "blocks in flight (jampack.cpp:215-219, :313-317); single-block latency under single_stream",
            "block_bytes": n, "units": 120,
            "l2": "per-block working set 6N = %d MB %s the 126 MB L2; no explicit flush" % (6 * n // 10**6, "exceeds" if 6 * n > 126e6 else "fits in")}


# ---- the other BASELINE.json configs, one line each (N = 1 only) -----------------------------------------------
def other_configs(jp, synth, torch, dev, with_cpu):
    """configs[2]: forward of 64 MiB uniform / repetitive / all-a blocks (and their inverse); configs[4]: one 256 MiB block,
    forward + inverse with the workspace each direction holds; configs[0]: the reference CLI's 8 MiB CPU round trip.
    One block at a time, resident in HBM, best of 3 warm runs; the CPU figure beside each is the unmodified reference,
    one block per host core (jampack.cpp:215-219 shape), one batch."""
    gold = {}
    try:
        for c in json.load(open(os.path.join(ROOT, "tests", "golden", "kat.json"))).get("big", []):
            gold[(c["kind"], c["len"], c["seed"])] = c["fnv_all"]
    except Exception:  # noqa: BLE001
        pass
    cores = os.cpu_count() or 1
    rows = []
    for kind, mib, seed, cfg in (("uniform", 64, 2, "configs[2]"), ("repetitive", 64, 3, "configs[2]"), ("alla", 64, 0, "configs[2]"), ("markov2", 256, 5, "configs[4]")):
        n = mib * MiB
        T = synth.gen(kind, n, seed)
        d_T = torch.from_numpy(T).to(dev)
        d_B = torch.zeros(n + TRAILER, dtype=torch.uint8, device=dev)
        d_back = torch.zeros(n, dtype=torch.uint8, device=dev)
        bf = bi = None
        for _ in range(4):
            jp.forward_device(d_T, d_B); s = jp.last_stats().asdict()
            if bf is None or s["ms_total"] < bf["ms_total"]:
                bf = s
        B = d_B.cpu().numpy()
        for _ in range(4):
            jp.inverse_device(d_B, d_back); s = jp.last_stats().asdict()
            if bi is None or s["ms_total"] < bi["ms_total"]:
                bi = s
        row = {"config": cfg, "input": f"{kind}({mib} MiB, seed {seed})", "forward_MBps": round(n / bf["ms_total"] / 1e3, 1), "forward_ms": round(bf["ms_total"], 3),
               "rounds": bf["rounds"], "sum_active_fraction": round(sum(bf["active_fraction"]), 4), "initial_depth": bf["initial_depth"],
               "run_bypass_suffixes": bf["bypass_suffixes"], "period": bf["period"],
               "forward_workspace_bytes_per_byte": round(bf["device_bytes"] / n, 2), "forward_phases_ms": bf["ms_phase"][:5],
               "inverse_MBps": round(n / bi["ms_total"] / 1e3, 1), "inverse_ms": round(bi["ms_total"], 3),
               "inverse_workspace_bytes_per_byte": round(bi["device_bytes"] / n, 3),
               "round_trip": bool(torch.equal(d_back, d_T)),
               "forward_matches_reference_hash": (("%016x" % synth.fnv(B)) == gold[(kind, n, seed)]) if (kind, n, seed) in gold else None}
        del d_T, d_B, d_back
        if with_cpu and mib <= 64:
            cf = cpu_reference("forward", T, B, cores)
            ci = cpu_reference("inverse", T, B, cores)
            row["cpu_baseline"] = {"forward_MBps": cf["value"], "inverse_MBps": ci["value"], "cores": cores, "kind": cf["kind"], "sample": cf["sample"],
                                   "output_matches": bool(cf["output_matches"] and ci["output_matches"])}
            row["forward_x_cpu"] = round(row["forward_MBps"] / cf["value"], 1)
            row["inverse_x_cpu"] = round(row["inverse_MBps"] / ci["value"], 1)
        rows.append(row)
    out = {"blocks": rows}
    ref_cli = os.path.join(ROOT, "oracle", "_ref", "Jampack_ref")
    if with_cpu and os.path.isfile(ref_cli):
        import hashlib
        import tempfile
        with tempfile.TemporaryDirectory() as d:
            src, jam, back = os.path.join(d, "in"), os.path.join(d, "out.jam"), os.path.join(d, "back")
            synth.gen("markov2", 8 * MiB, 1).tofile(src)
            t0 = time.perf_counter(); subprocess.run([ref_cli, "c", src, jam], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=False); tc = time.perf_counter() - t0
            t0 = time.perf_counter(); subprocess.run([ref_cli, "d", jam, back], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=False); td = time.perf_counter() - t0
            sha = lambda p: hashlib.sha256(open(p, "rb").read()).hexdigest() if os.path.isfile(p) else None  # noqa: E731
            out["config0_cpu_round_trip"] = {"input": "markov2(8 MiB, seed 1)", "command": "Jampack_ref c / d, default settings (reference main.cpp)",
                                             "compress_s": round(tc, 2), "decompress_s": round(td, 2), "jam_bytes": os.path.getsize(jam) if os.path.isfile(jam) else None,
                                             "jam_sha256": sha(jam), "round_trip": sha(back) == sha(src), "threads": cores}
    return out


# ---- our arm ----------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch
    import torch.distributed as dist
    import jampack_b200 as jp
    import synth
    from jampack_b200 import build, shard

    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: this stage has no CPU path"}))
        return 2
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner there at NCCL_DEBUG=VERSION and WARN
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG", None)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    build.build()
    jp.set_devices([local])

    n = args.block_mib * MiB
    nlen = n - n % 120
    T = synth.gen(args.kind, n, args.seed)
    K, W = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # -- inputs resident in HBM
    d_T = torch.from_numpy(T).to(dev)
    d_B = torch.zeros(n + TRAILER, dtype=torch.uint8, device=dev)
    d_back = torch.zeros(n, dtype=torch.uint8, device=dev)
    jp.forward_device(d_T, d_B)
    B = d_B.cpu().numpy()
    parity = {"round_trip": None, "forward_matches_golden": None}
    try:
        gold = json.load(open(os.path.join(ROOT, "tests", "golden", "kat.json")))
        for c in gold.get("big", []):
            if (c["kind"], c["len"], c["seed"]) == (args.kind, n, args.seed):
                parity["forward_matches_golden"] = ("%016x" % synth.fnv(B)) == c["fnv_all"]
    except Exception:  # noqa: BLE001
        pass

    def timed_device(fn, stat_keys):
        for _ in range(W):
            fn()
        acc = {k: 0.0 for k in stat_keys}
        launches = 0
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            fn()
            st = jp.last_stats()
            launches += st.kernel_launches
            for k in stat_keys:
                acc[k] += st.ms_phase[k]
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        return ms, acc, launches, jp.last_stats().asdict()

    def timed_host(fn):
        for _ in range(W):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            fn()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        barrier()
        return ms

    # The reference drives its stage from an OpenMP team, one block per worker (jampack.cpp:215-219, :313-317), so
    # several blocks are in flight per device and their copies overlap other blocks' kernels. `callers` host
    # threads each push K blocks through the host C-ABI from their own pinned buffers.
    def timed_host_concurrent(direction, callers):
        src = B if direction == "inverse" else T
        bufs = []
        for _ in range(callers):
            hi, ho = jp.PinnedBlock(n + TRAILER), jp.PinnedBlock(n + TRAILER)
            hi.array[: src.size] = src
            bufs.append((hi, ho))
        call = (lambda hi, ho: jp.inverse(hi.array, out=ho.array)) if direction == "inverse" else \
               (lambda hi, ho: jp.forward(hi.array[:n], out=ho.array))

        def worker(i, reps):
            for _ in range(reps):
                call(*bufs[i])

        def run(reps):
            th = [threading.Thread(target=worker, args=(i, reps)) for i in range(callers)]
            [t.start() for t in th]
            [t.join() for t in th]

        run(max(W, 1))
        barrier()
        t0 = time.perf_counter()
        run(K)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        barrier()
        want = T if direction == "inverse" else B
        ok = all(bool((ho.array[: want.size] == want).all()) for _, ho in bufs)
        for hi, ho in bufs:
            hi.free(); ho.free()
        return ms, ok

    # The same callers doing ONLY the copies of those calls (jp_bwt_debug_copy: pinned block -> device, device -> pinned
    # block, no kernels): what the host side of this box allows, i.e. the ceiling of the end-to-end figure.
    def timed_copy_ceiling(direction, callers):
        nin, nout = (n + TRAILER, n) if direction == "inverse" else (n, n + TRAILER)
        bufs = [(jp.PinnedBlock(nin), jp.PinnedBlock(nout)) for _ in range(callers)]

        def worker(i, reps):
            for _ in range(reps):
                jp.debug_copy(bufs[i][0].array, bufs[i][1].array)

        def run(reps):
            th = [threading.Thread(target=worker, args=(i, reps)) for i in range(callers)]
            [t.start() for t in th]
            [t.join() for t in th]

        run(max(W, 1))
        barrier()
        t0 = time.perf_counter()
        run(K)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        barrier()
        for hi, ho in bufs:
            hi.free(); ho.free()
        return ms

    # Pageable caller blocks, as the unmodified reference allocates them (calloc, jampack.cpp:74-76): the library
    # page-locks a block in the background after its first call, so the first call is a staged copy and the timed ones DMA directly.
    def timed_host_pageable(direction, callers):
        src = B if direction == "inverse" else T
        bufs = [(np.array(src, copy=True), np.zeros(n + TRAILER, dtype=np.uint8)) for _ in range(callers)]
        call = (lambda hi, ho: jp.inverse(hi, out=ho)) if direction == "inverse" else (lambda hi, ho: jp.forward(hi, out=ho))

        def worker(i, reps):
            for _ in range(reps):
                call(*bufs[i])

        def run(reps):
            th = [threading.Thread(target=worker, args=(i, reps)) for i in range(callers)]
            [t.start() for t in th]
            [t.join() for t in th]

        jp.keep_host_blocks_locked(True)               # these blocks live for the whole leg, like the reference's
        t0 = time.perf_counter()
        run(1)
        first_ms = (time.perf_counter() - t0) * 1e3
        time.sleep(0.5)                                # the caller's other stages: the background page-locking happens here
        run(max(W - 1, 1))
        barrier()
        t0 = time.perf_counter()
        run(K)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        barrier()
        want = T if direction == "inverse" else B
        ok = all(bool((ho[: want.size] == want).all()) for _, ho in bufs)
        jp.host_release()                              # before the blocks are freed
        jp.keep_host_blocks_locked(False)
        return ms, ok, first_ms

    # Device-resident throughput with `callers` blocks in flight on their own streams (same reason as above: a
    # block's latency-bound tails -- the end of each LF walk, the single-block scans -- overlap another block's work).
    def timed_device_concurrent(direction, callers):
        ins = [(d_B if direction == "inverse" else d_T).clone() for _ in range(callers)]
        outs = [torch.zeros(n + TRAILER, dtype=torch.uint8, device=dev) for _ in range(callers)]
        streams = [torch.cuda.Stream(device=dev) for _ in range(callers)]
        launches = [0] * callers

        def worker(i, reps):
            torch.cuda.set_device(local)
            with torch.cuda.stream(streams[i]):
                for _ in range(reps):
                    if direction == "inverse":
                        jp.inverse_device(ins[i], outs[i])
                    else:
                        jp.forward_device(ins[i], outs[i])
                    launches[i] += jp.last_stats().kernel_launches

        def run(reps):
            th = [threading.Thread(target=worker, args=(i, reps)) for i in range(callers)]
            [t.start() for t in th]
            [t.join() for t in th]

        run(max(W, 1))
        launches[:] = [0] * callers
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(K)                                         # every call returns after its own stream has drained
        e1.record()
        barrier()
        want = d_T if direction == "inverse" else d_B
        ok = all(bool(torch.equal(o[: want.numel()], want)) for o in outs)
        return e0.elapsed_time(e1), ok, sum(launches)

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()

    # -- inverse: device-resident, then end to end through the host C-ABI with pinned blocks
    inv_ms, inv_acc, inv_launches, inv_stats = timed_device(lambda: jp.inverse_device(d_B, d_back), [0, 1, 2, 3, 4])
    parity["round_trip"] = bool(torch.equal(d_back, d_T))
    h_in, h_out = jp.PinnedBlock(n + TRAILER), jp.PinnedBlock(n + TRAILER)
    h_in.array[:] = B
    inv_e2e_ms = timed_host(lambda: jp.inverse(h_in.array, out=h_out.array))
    parity["round_trip"] = parity["round_trip"] and bool((h_out.array[:n] == T).all())
    CALLERS = args.callers
    inv_e2e_c_ms, okc = timed_host_concurrent("inverse", CALLERS)
    parity["round_trip"] = parity["round_trip"] and okc
    inv_c_ms, okc, inv_c_launches = timed_device_concurrent("inverse", CALLERS)
    parity["round_trip"] = parity["round_trip"] and okc
    inv_copy_ms = timed_copy_ceiling("inverse", CALLERS)
    inv_page_ms, okc, inv_page_first_ms = timed_host_pageable("inverse", CALLERS)
    parity["round_trip"] = parity["round_trip"] and okc

    fwd = None
    if not args.no_forward:
        fwd_ms, fwd_acc, fwd_launches, fwd_stats = timed_device(lambda: jp.forward_device(d_T, d_B), [0, 1, 2, 3, 4])
        h_in.array[:n] = T
        fwd_e2e_ms = timed_host(lambda: jp.forward(h_in.array[:n], out=h_out.array))
        parity["forward_host_equals_device"] = bool((h_out.array[: n + TRAILER] == B).all())
        fwd_e2e_c_ms, okc = timed_host_concurrent("forward", CALLERS)
        parity["forward_host_equals_device"] = parity["forward_host_equals_device"] and okc
        fwd_c_ms, okc, fwd_c_launches = timed_device_concurrent("forward", CALLERS)
        parity["forward_host_equals_device"] = parity["forward_host_equals_device"] and okc
        fwd_copy_ms = timed_copy_ceiling("forward", CALLERS)
        fwd_page_ms, okc, _ = timed_host_pageable("forward", CALLERS)
        parity["forward_host_equals_device"] = parity["forward_host_equals_device"] and okc
        fwd = (fwd_ms, fwd_acc, fwd_launches, fwd_stats, fwd_e2e_ms, fwd_e2e_c_ms, fwd_c_ms, fwd_c_launches, fwd_copy_ms, fwd_page_ms)

    clk = clocks.stop() if rank == 0 else None

    inv_ms_max, total_bytes = shard.reduce_step_stats(inv_ms, n * K, dev)
    inv_e2e_max, _ = shard.reduce_step_stats(inv_e2e_ms, n * K, dev)
    inv_e2e_c_max, _ = shard.reduce_step_stats(inv_e2e_c_ms, n * K, dev)
    inv_c_max, _ = shard.reduce_step_stats(inv_c_ms, n * K, dev)
    inv_copy_max, _ = shard.reduce_step_stats(inv_copy_ms, n * K, dev)
    inv_page_max, _ = shard.reduce_step_stats(inv_page_ms, n * K, dev)
    if fwd:
        fwd_ms_max, _ = shard.reduce_step_stats(fwd[0], n * K, dev)
        fwd_e2e_max, _ = shard.reduce_step_stats(fwd[4], n * K, dev)
        fwd_e2e_c_max, _ = shard.reduce_step_stats(fwd[5], n * K, dev)
        fwd_c_max, _ = shard.reduce_step_stats(fwd[6], n * K, dev)
        fwd_copy_max, _ = shard.reduce_step_stats(fwd[8], n * K, dev)
        fwd_page_max, _ = shard.reduce_step_stats(fwd[9], n * K, dev)

    if rank == 0:
        peak, peak_src = measured_peaks()
        # random 4-byte gathers over a table the size of this block's LF table (4*nlen bytes, rounded up to 2^k)
        rand_rate = jp.debug_gather_rate(1 << (4 * nlen - 1).bit_length(), 148 * 2048, 256, True)
        walk_ms = (inv_acc[2] + inv_acc[3] + inv_acc[4]) / K                        # both walk kernels + ranking
        algo_bytes = 64.0 * nlen                                                    # 2 random sectors / byte
        achieved = algo_bytes / (walk_ms * 1e-3) / 1e9
        single_walk = inv_stats.get("stream_chunks", 0) > 0       # blocks of 30 Mi and more: every LF entry gathered once
        walk_kernels = "k_inv_walk_stream + k_inv_rank_packed + k_inv_place" if single_walk else "k_inv_walk_len + k_inv_rank + k_inv_walk_emit"
        phase_names = ("hist_ctable", "lf_build", "walk_stream", "rank", "place") if single_walk else ("hist_ctable", "lf_build", "walk_len", "rank", "walk_emit")
        roof = {"bound": "hbm", "kernel": walk_kernels, "achieved": round(achieved, 1),
                "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "peak_source": peak_src,
                "algorithmic_bytes_per_launch_pair": algo_bytes, "ms_per_step": round(walk_ms, 4),
                "traffic": ncu_traffic("inverse_walk_single" if single_walk else "inverse_walk"),
                "rand_peak": round(rand_rate * 32 / 1e9, 1), "rand_unit": "GB/s of 32 B sectors = live dependent-gather micro-benchmark over a table of the LF table's size x 32 B",
                "rand_gathers_per_s": round(rand_rate / 1e9, 2),
                "frac_of_rand": round(achieved / (rand_rate * 32 / 1e9), 4),
                "phases_ms": {k: round(v / K, 4) for k, v in zip(phase_names, inv_acc.values())}}
        line = {"metric": "inv BWT MB/s", "value": round(CALLERS * total_bytes / (inv_c_max * 1e-3) / 1e6, 1), "unit": "MB/s",
                "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": round(inv_c_max / K, 4), "blocks_per_step_per_gpu": CALLERS,
                "single_stream": {"value": round(total_bytes / (inv_ms_max * 1e-3) / 1e6, 1), "ms_per_block": round(inv_ms_max / K, 4),
                                  "note": "one block at a time on one stream: the latency view the roofline phases below are taken from"},
                "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(args, n),
                "e2e": {"value": round(CALLERS * total_bytes / (inv_e2e_c_max * 1e-3) / 1e6, 1), "unit": "MB/s",
                        "ms_per_step": round(inv_e2e_c_max / K, 4), "blocks_per_step": CALLERS,
                        "mode": f"{CALLERS} concurrent host callers per GPU, one block each per step, through jp_bwt_inverse (how the "
                                "reference's OpenMP block loop drives the stage); copies of one block overlap kernels of another",
                        "h2d_bytes_per_step": CALLERS * (n + TRAILER), "d2h_bytes_per_step": CALLERS * n, "host_memory": "pinned",
                        "copy_ceiling": {"value": round(CALLERS * total_bytes / (inv_copy_max * 1e-3) / 1e6, 1), "unit": "MB/s",
                                         "what": "the same callers, blocks and byte counts through jp_bwt_debug_copy: pinned H2D + D2H only, no kernels"},
                        "frac_of_copy_ceiling": round(inv_copy_max / inv_e2e_c_max, 4),
                        "pageable": {"value": round(CALLERS * total_bytes / (inv_page_max * 1e-3) / 1e6, 1), "unit": "MB/s",
                                     "first_call_ms": round(inv_page_first_ms, 2),
                                     "what": "the same callers with pageable (malloc) blocks, as the unmodified reference allocates them; the library "
                                             "page-locks a block in the background after its first call (first_call_ms = that first, staged call), later calls DMA directly"},
                        "single_caller": {"value": round(total_bytes / (inv_e2e_max * 1e-3) / 1e6, 1), "ms_per_step": round(inv_e2e_max / K, 4),
                                          "h2d_bytes_per_step": n + TRAILER, "d2h_bytes_per_step": n}},
                "gpu_launches": inv_c_launches, "clocks": clk, "roofline": roof, "parity": parity,
                "inverse_stats": {k: inv_stats[k] for k in ("subchains", "subchain_spacing", "device_bytes", "kernel_launches", "stream_chunks", "random_sectors")},
                "host_cores": os.cpu_count()}
        if fwd:
            st = fwd[3]
            sum_a = sum(st["active_fraction"])
            a_fwd = nlen * (38 + 24) + 80.0 * nlen * sum_a                          # SURVEY.md 8d counted model
            r_fwd = nlen * 32 + 64.0 * nlen * sum_a
            f_ms = fwd_ms_max / K
            line["forward"] = {"value": round(CALLERS * total_bytes / (fwd_c_max * 1e-3) / 1e6, 1), "unit": "MB/s", "ms_per_step": round(fwd_c_max / K, 4),
                               "blocks_per_step_per_gpu": CALLERS,
                               "single_stream": {"value": round(total_bytes / (fwd_ms_max * 1e-3) / 1e6, 1), "ms_per_block": round(f_ms, 4)},
                               "e2e": {"value": round(CALLERS * total_bytes / (fwd_e2e_c_max * 1e-3) / 1e6, 1), "unit": "MB/s",
                                       "blocks_per_step": CALLERS, "h2d_bytes_per_step": CALLERS * n, "d2h_bytes_per_step": CALLERS * (n + TRAILER),
                                       "host_memory": "pinned",
                                       "copy_ceiling": {"value": round(CALLERS * total_bytes / (fwd_copy_max * 1e-3) / 1e6, 1), "unit": "MB/s"},
                                       "frac_of_copy_ceiling": round(fwd_copy_max / fwd_e2e_c_max, 4),
                                       "pageable": {"value": round(CALLERS * total_bytes / (fwd_page_max * 1e-3) / 1e6, 1), "unit": "MB/s"},
                                       "single_caller": {"value": round(total_bytes / (fwd_e2e_max * 1e-3) / 1e6, 1), "ms_per_step": round(fwd_e2e_max / K, 4)}},
                               "gpu_launches": fwd[7], "rounds": st["rounds"], "active_fraction": st["active_fraction"],
                               "symbol_bits": st["symbol_bits"], "initial_depth": st["initial_depth"], "device_bytes": st["device_bytes"],
                               "workspace_bytes_per_byte": round(st["device_bytes"] / n, 2),
                               "roofline": {"bound": "hbm", "achieved": round(a_fwd / (f_ms * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                                            "frac": round(a_fwd / (f_ms * 1e-3) / 1e9 / peak, 4), "algorithmic_bytes": a_fwd,
                                            "random_bytes": r_fwd, "traffic": ncu_traffic("forward")},
                               "phases_ms": {k: round(v / K, 4) for k, v in zip(("keys", "initial_sort", "initial_ranks", "rounds", "emit"), fwd[1].values())}}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            cb = cpu_reference("inverse", T, B, cores)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "output_matches")}
            line["cpu_baseline"]["single_block"] = cpu_reference_single_block("inverse", T, B, cores)   # the other all-core shape; `value` is the better one
            sb = line["cpu_baseline"]["single_block"]
            if sb and sb["value"] > line["cpu_baseline"]["value"]:
                line["cpu_baseline"].update(value=sb["value"], sample=sb["sample"])
            if fwd:
                cf = cpu_reference("forward", T, B, cores)
                line["forward"]["cpu_baseline"] = {k: cf[k] for k in ("value", "unit", "cores", "kind", "sample", "output_matches")}
                line["forward"]["cpu_baseline"]["single_block"] = cpu_reference_single_block("forward", T, B, cores)
                sb = line["forward"]["cpu_baseline"]["single_block"]
                if sb and sb["value"] > line["forward"]["cpu_baseline"]["value"]:
                    line["forward"]["cpu_baseline"].update(value=sb["value"], sample=sb["sample"])
            line["legacy_cuda_baseline"] = legacy_cuda_baseline(T, B)
        if world == 1 and not args.no_configs:
            line["configs"] = other_configs(jp, synth, torch, dev, not args.no_cpu_baseline)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
