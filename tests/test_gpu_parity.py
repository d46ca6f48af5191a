"""GPU parity tests (run on the B200 box: pytest -m gpu). Everything goes through the C-ABI of include/jp_bwt.h
(ctypes, jampack_b200/__init__.py) and is compared bit-for-bit with the oracle: the compiled reference
(oracle/_ref) when it travelled with the repo, else the C restatement; plus the committed golden vectors."""
import hashlib
import os
import subprocess
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
MiB = 1 << 20
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _impl(orc):
    return "ref" if orc.ref() is not None else "port"


SMALL = [("kat_quadratic", 240, 0), ("kat_quadratic", 250, 0), ("alla", 360, 0), ("kat_extremes", 240, 0),
         ("kat_quadratic", 120, 0), ("uniform", 121, 7), ("markov2", 4093, 9), ("repetitive", 70000, 3),
         ("uniform", 5000, 4), ("markov2", 65536 + 120, 3), ("uniform", 4096 * 3 + 1, 5)]
MEDIUM = [("markov2", MiB, 1), ("uniform", MiB, 2), ("repetitive", MiB, 3), ("alla", MiB, 0), ("markov2", 3 * MiB + 77, 4)]


@pytest.mark.parametrize("kind,n,seed", SMALL + MEDIUM)
def test_forward_bit_exact(jp, orc, kind, n, seed):
    T = orc.gen(kind, n, seed)
    want = orc.forward(T, _impl(orc), prefill=0x5C)
    got = jp.forward(T, prefill=0x5C)
    assert got.size == n + 480
    assert (got == want).all()


@pytest.mark.parametrize("kind,n,seed", SMALL + MEDIUM)
def test_inverse_bit_exact_and_cross(jp, orc, kind, n, seed):
    T = orc.gen(kind, n, seed)
    B = orc.forward(T, _impl(orc))
    assert (jp.inverse(B) == T).all()                       # reference forward -> our inverse
    ours = jp.forward(T)
    assert (orc.inverse(ours, _impl(orc), threads=4) == T).all()   # our forward -> reference inverse


@pytest.mark.parametrize("n", [0, 1, 2, 119])
def test_short_blocks_copy_and_leave_trailer(jp, orc, n):
    """len < 120: nothing is transformed and the 480 trailer bytes are NOT written (bwt.cpp:35, SURVEY finding 7)."""
    T = orc.gen("uniform", n, 3)
    out = jp.forward(T, prefill=0xAB)
    assert (out[:n] == T).all() and (out[n:] == 0xAB).all()
    assert (out == orc.forward(T, _impl(orc), prefill=0xAB)).all()
    back = jp.inverse(out)
    assert back.size == n and (back == T).all()


def test_zero_runs_and_extreme_bytes(jp, orc):
    """0x00 runs at the end are where zero-padded prefix keys would tie with the end of string."""
    T = orc.gen("markov2", 120 * 300, 8)
    T[-500:] = 0
    T[1000:1400] = 0
    T[2000:2100] = 255
    want = orc.forward(T, _impl(orc))
    assert (jp.forward(T) == want).all()
    assert (jp.inverse(want) == T).all()
    Z = np.zeros(120 * 50, dtype=np.uint8)
    assert (jp.forward(Z) == orc.forward(Z, _impl(orc))).all()
    F = np.full(120 * 50 + 3, 255, dtype=np.uint8)
    assert (jp.forward(F) == orc.forward(F, _impl(orc))).all()


def test_golden_vectors(jp, orc, golden):
    """tests/golden/kat.json: outputs of the unmodified reference (incl. SURVEY.md Appendix B indices)."""
    for c in golden["cases"]:
        T = orc.gen(c["kind"], c["len"], c["seed"])
        got = jp.forward(T)
        if "out_hex" in c:
            want = np.frombuffer(bytes.fromhex(c["out_hex"]), dtype=np.uint8)
            assert (got[: want.size] == want).all(), c["name"]
        assert "%016x" % orc.fnv(got[: c["len"]]) == c["fnv_bwt"], c["name"]
        if c["nlen"]:
            assert "%016x" % orc.fnv(got) == c["fnv_all"], c["name"]
            assert [int(x) for x in orc.indices(got)] == c["indices"], c["name"]
            assert (jp.inverse(got) == T).all(), c["name"]


def test_golden_full_size_blocks(jp, orc, golden):
    """BASELINE.json configs 2, 3 and 5 at full size (64 MiB x4, 256 MiB): forward hash + all 120 indices
    against the reference's, and the round trip."""
    assert len(golden.get("big", [])) == 5
    for c in golden["big"]:
        T = orc.gen(c["kind"], c["len"], c["seed"])
        assert "%016x" % orc.fnv(T) == c["fnv_in"]
        got = jp.forward(T)
        st = jp.last_stats()
        assert "%016x" % orc.fnv(got[: c["len"]]) == c["fnv_bwt"], c["name"]
        assert "%016x" % orc.fnv(got) == c["fnv_all"], c["name"]
        assert [int(x) for x in orc.indices(got)] == c["indices"], c["name"]
        back = jp.inverse(got)
        si = jp.last_stats()
        assert (back == T).all(), c["name"]
        # inverse stays within 6N + o(N): in (N+480) + out (N) are the caller's, the workspace is lf (4N) + side tables
        assert si.device_bytes <= 4 * c["nlen"] + c["nlen"] // 4 + (1 << 20), (c["name"], si.device_bytes)
        # blocks this large decode in one walk (every LF entry gathered once); the stream stays inside in + out
        assert si.stream_chunks > 0 and si.stream_chunks * 1024 <= c["nlen"] + c["nlen"] // 2, (c["name"], si.stream_chunks)
        assert st.rounds <= 40
        # forward workspace: six units of 4(N+2) bytes + side tables = 24.3 N (+ 4.1 N of repeat lengths on a periodic block,
        # + the call's own copies of the block and of the output when it comes from host memory); it was 46.3 N in round 1
        assert st.device_bytes <= (31 if st.period else 27) * c["nlen"] + (16 << 20), (c["name"], st.device_bytes / c["nlen"])


def test_lf_table_is_inverse_of_reference_map(jp, orc):
    """Rows a7/a8 of SURVEY.md 8a: C table == bwt.cpp:141-169, LF == inverse permutation of Map (bwt.cpp:171-174)."""
    for kind, n, seed in [("markov2", 120 * 1000, 1), ("uniform", 65536 * 2 + 120 * 5, 2), ("alla", 12000, 0)]:
        T = orc.gen(kind, n, seed)
        B = orc.forward(T, _impl(orc))
        nlen = n - n % 120
        idx = int(orc.indices(B)[0])
        Map, Ct = orc.build_map(B[:nlen], nlen, idx)
        lf, ct = jp.debug_lf(B[:nlen])
        assert (ct == Ct).all()
        i = np.arange(nlen)
        assert (Map[lf - 1] == i + (i >= idx)).all()


def test_suffix_array_contract(jp, orc):
    """Row a2: the suffix array itself (divsufsort contract), checked by brute force on a small block."""
    rng = np.random.default_rng(5)
    T = rng.integers(0, 3, 3000).astype(np.uint8)
    T[1500:] = T[:1500]
    sa = jp.debug_suffix_array(T)
    b = T.tobytes()
    assert sa.tolist() == sorted(range(T.size), key=lambda i: b[i:])
    T2 = orc.gen("markov2", 200000, 12)
    assert (jp.debug_suffix_array(T2) == orc.suffix_array(T2)).all()


def test_stage_interface_mirror(jp, orc):
    """BlockSort::Bwt through Buffer/Options, incl. the size side effects (bwt.cpp:27, :77-78)."""
    n = 120 * 777 + 13
    T = orc.gen("markov2", n, 21)
    cap = int(n * 1.05) + 480
    Input = jp.Buffer(np.zeros(cap, dtype=np.uint8), n)
    Input.block[:n] = T
    Output = jp.Buffer(np.zeros(cap, dtype=np.uint8), 0)
    bwt = jp.Bwt()
    bwt.ForwardBwt(Input, Output)
    assert Output.size[0] == n + 480 and Input.size[0] == n
    assert (Output.block[: n + 480] == orc.forward(T, _impl(orc))).all()
    Back = jp.Buffer(np.zeros(cap, dtype=np.uint8), 0)
    bwt.InverseBwt(Output, Back, jp.Options(Threads=8))
    assert Output.size[0] == n and Back.size[0] == n           # InverseBwt shrinks *Input.size in place
    assert (Back.block[:n] == T).all()


def test_inverse_rejects_corrupt_input(jp, orc):
    T = orc.gen("markov2", 120 * 2000, 5)
    n = T.size
    B = orc.forward(T, _impl(orc)).copy()
    bad = B.copy()
    bad[n + 4 * 7: n + 4 * 7 + 4] = np.frombuffer(np.int32(n + 1).tobytes(), dtype=np.uint8)   # out of range
    with pytest.raises(jp.BwtError) as e:
        jp.inverse(bad)
    assert e.value.rc == -5
    bad = B.copy()
    bad[n + 4 * 9: n + 4 * 9 + 4] = bad[n + 4 * 8: n + 4 * 8 + 4]                                # duplicated index
    with pytest.raises(jp.BwtError):
        jp.inverse(bad)
    bad = B.copy()
    bad[n + 4 * 50: n + 4 * 50 + 4] = np.frombuffer(np.int32(12345).tobytes(), dtype=np.uint8)   # wrong but in range
    with pytest.raises(jp.BwtError) as e:
        jp.inverse(bad)
    assert e.value.rc == -6
    assert (jp.inverse(B) == T).all()                           # the context is still healthy afterwards


def test_device_resident_entry_points(jp, orc):
    import torch
    T = orc.gen("markov2", 4 * MiB + 50, 31)
    want = orc.forward(T, _impl(orc))
    d_in = torch.from_numpy(T).cuda()
    d_out = jp.forward_device(d_in)
    st = jp.last_stats()
    assert st.kernel_launches > 0 and st.ms_total > 0
    torch.cuda.synchronize()
    assert (d_out.cpu().numpy() == want).all()
    d_back = jp.inverse_device(d_out)
    assert (d_back.cpu().numpy()[: T.size] == T).all()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        d_back2 = jp.inverse_device(d_out)
    s.synchronize()
    assert torch.equal(d_back, d_back2)


def test_pinned_host_blocks(jp, orc):
    n = 2 * MiB
    T = orc.gen("uniform", n, 77)
    src = jp.PinnedBlock(n + 480)
    dst = jp.PinnedBlock(n + 480)
    src.array[:n] = T
    out = jp.forward(src.array[:n], out=dst.array)
    assert (out == orc.forward(T, _impl(orc))).all()
    src.free(); dst.free()


def test_concurrent_calls_are_reentrant(jp, orc):
    """The reference calls its stage from an OpenMP team, one block per thread (jampack.cpp:215-219)."""
    blocks = [orc.gen("markov2", MiB + 120 * i, 40 + i) for i in range(8)]
    wants = [orc.forward(b, _impl(orc)) for b in blocks]
    outs = [None] * len(blocks)
    errs = []

    def work(i):
        try:
            f = jp.forward(blocks[i])
            outs[i] = (f, jp.inverse(f))
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(blocks))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for i, b in enumerate(blocks):
        assert (outs[i][0] == wants[i]).all() and (outs[i][1] == b).all()


def test_linearity_free_properties_at_scale(jp, orc):
    """Size-independent properties on a block nobody has a golden hash for: the BWT is a permutation of the
    block's bytes, the indices are distinct rows in [1, nlen], and inverse(forward(x)) == x."""
    n = 96 * MiB + 119
    T = orc.gen("markov2", n, 4242)
    out = jp.forward(T)
    nlen = n - n % 120
    assert (np.bincount(out[:nlen], minlength=256) == np.bincount(T[:nlen], minlength=256)).all()
    assert (out[nlen:n] == T[nlen:]).all()
    I = orc.indices(out)
    assert I.min() >= 1 and I.max() <= nlen and np.unique(I).size == 120
    assert (jp.inverse(out) == T).all()


def test_reference_pipeline_with_shim_is_byte_identical(jp, orc, golden, tmp_path):
    """BASELINE.json configs[0]: the reference CLI with our stage linked in place of bwt.cpp must emit the same
    .jam bytes as the unmodified reference (sha256 pinned in tests/golden/kat.json), and each must decode the other's."""
    shim = os.path.join(ROOT, "oracle", "_ref", "Jampack_shim")
    ref = os.path.join(ROOT, "oracle", "_ref", "Jampack_ref")
    if not (os.path.isfile(shim) and os.path.isfile(ref)):
        pytest.skip("oracle/_ref binaries did not travel (built only where /root/reference exists)")
    T = orc.gen("markov2", 8 * MiB, 1)
    src = tmp_path / "in.bin"
    T.tofile(src)
    assert hashlib.sha256(T.tobytes()).hexdigest() == golden["config1"]["input_sha256"]
    jam_s, jam_r = tmp_path / "shim.jam", tmp_path / "ref.jam"
    subprocess.run([shim, "c", str(src), str(jam_s), "-t4"], check=True, stdout=subprocess.DEVNULL, timeout=600)
    sha = hashlib.sha256(jam_s.read_bytes()).hexdigest()
    assert jam_s.stat().st_size == golden["config1"]["jam_bytes"]
    assert sha == golden["config1"]["jam_sha256"]
    subprocess.run([ref, "c", str(src), str(jam_r), "-t4"], check=True, stdout=subprocess.DEVNULL, timeout=600)
    assert jam_r.read_bytes() == jam_s.read_bytes()
    back_s, back_r = tmp_path / "back_s.bin", tmp_path / "back_r.bin"
    subprocess.run([shim, "d", str(jam_r), str(back_s), "-t4"], check=True, stdout=subprocess.DEVNULL, timeout=600)
    subprocess.run([ref, "d", str(jam_s), str(back_r), "-t4"], check=True, stdout=subprocess.DEVNULL, timeout=600)
    assert back_s.read_bytes() == T.tobytes() and back_r.read_bytes() == T.tobytes()
    # single-block decode mode (-T, jampack.cpp:249-282) goes through the same stage from the main thread
    back_t = tmp_path / "back_t.bin"
    subprocess.run([shim, "d", str(jam_s), str(back_t), "-T"], check=True, stdout=subprocess.DEVNULL, timeout=600)
    assert back_t.read_bytes() == T.tobytes()


# ---- widening: the suffix sorter under the reference's second divsufsort call site ------
@pytest.mark.parametrize("kind,n,seed", [("kat_extremes", 1, 0), ("kat_extremes", 2, 0), ("uniform", 3, 1), ("alla", 1000, 0),
                                         ("markov2", 100003, 2), ("repetitive", 300001, 3), ("uniform", MiB + 1, 4),
                                         ("markov2", 2 * MiB, 6)])
def test_suffix_array_equals_divsufsort(jp, orc, kind, n, seed):
    """jp_bwt_suffix_array == divsufsort(T, SA, n) (divsufsort.cpp:1721) for any length, not only multiples of 120."""
    T = orc.gen(kind, n, seed)
    want = orc.suffix_array(T, "ref") if orc.ref() is not None else orc.suffix_array(T)
    got = jp.suffix_array(T)
    assert (got == want).all()
    assert jp.last_stats().kernel_launches > 0
    assert jp.suffix_array(np.zeros(0, dtype=np.uint8)).size == 0


def test_text_with_long_repeats_takes_the_large_group_route(jp, orc):
    """Natural-language-like input: short groups, groups that need the shared-memory radix kernel, and groups longer
    than a window (repeated paragraphs) in the same block -- all three refinement routes must agree with the oracle."""
    rng = np.random.default_rng(7)
    words = [bytes(rng.integers(97, 123, rng.integers(2, 9)).astype(np.uint8)) for _ in range(300)]
    para = b" ".join(words[i] for i in rng.integers(0, 300, 4000))
    parts = []
    for r in range(60):
        parts.append(para[: rng.integers(2000, len(para))])
        parts.append(b" ".join(words[i] for i in rng.integers(0, 300, 3000)))
        parts.append(b"=" * int(rng.integers(10, 9000)))
    T = np.frombuffer(b"".join(parts), dtype=np.uint8).copy()
    T = T[: T.size - T.size % 120 + 7]
    want = orc.forward(T, _impl(orc))
    got = jp.forward(T)                                     # as shipped: the runs of '=' are placed by the run bypass
    assert (got == want).all() and jp.last_stats().bypass_runs > 0
    saved = os.environ.get("JP_BWT_FWD_BYPASS")
    os.environ["JP_BWT_FWD_BYPASS"] = "0"                   # without it they are over-long groups: the large-group route
    try:
        got = jp.forward(T)
        st = jp.last_stats()
    finally:
        if saved is None:
            os.environ.pop("JP_BWT_FWD_BYPASS", None)
        else:
            os.environ["JP_BWT_FWD_BYPASS"] = saved
    assert (got == want).all()
    assert st.large_fraction > 0, "expected some suffixes on the large-group route"
    assert (jp.inverse(got) == T).all()


def test_reference_pipeline_m2_uses_gpu_suffix_sorter(jp, orc, tmp_path):
    """-m2 (lz77.cpp:134-146) with divsufsort() bound to jp_bwt_suffix_array: same .jam bytes as the reference."""
    shim = os.path.join(ROOT, "oracle", "_ref", "Jampack_shim_m2")
    ref = os.path.join(ROOT, "oracle", "_ref", "Jampack_ref")
    if not (os.path.isfile(shim) and os.path.isfile(ref)):
        pytest.skip("oracle/_ref binaries did not travel")
    T = orc.gen("markov2", MiB // 2, 77)          # the match finder itself is "incredibly slow" (lz77.cpp:134): keep it small
    src = tmp_path / "in.bin"
    T.tofile(src)
    jam_s, jam_r, back = tmp_path / "s.jam", tmp_path / "r.jam", tmp_path / "back.bin"
    subprocess.run([shim, "c", str(src), str(jam_s), "-m2", "-b1", "-t2"], check=True, stdout=subprocess.DEVNULL, timeout=900)
    subprocess.run([ref, "c", str(src), str(jam_r), "-m2", "-b1", "-t2"], check=True, stdout=subprocess.DEVNULL, timeout=900)
    assert jam_s.read_bytes() == jam_r.read_bytes()
    subprocess.run([shim, "d", str(jam_r), str(back), "-t2"], check=True, stdout=subprocess.DEVNULL, timeout=900)
    assert back.read_bytes() == T.tobytes()


def test_consume_variant_and_trace(jp, orc, tmp_path):
    """jp_bwt_inverse_device_consume gives the same block while keeping the workspace at the LF table + histograms
    (the records live in the dead input block); JP_BWT_TRACE prints the stage's wall-clock totals at exit."""
    import torch
    n = 8 * MiB
    T = orc.gen("markov2", n, 3)
    B = jp.forward(T)
    d_in = torch.from_numpy(B.copy()).cuda()
    out_a = jp.inverse_device(d_in)
    bytes_const = jp.last_stats().device_bytes
    out_b = jp.inverse_device(d_in.clone(), consume=True)
    bytes_consume = jp.last_stats().device_bytes
    assert torch.equal(out_a, out_b) and (out_a.cpu().numpy()[:n] == T).all()
    assert bytes_consume < bytes_const and bytes_consume <= 4 * n + n // 32 + (1 << 20)
    code = ("import numpy as np, jampack_b200 as jp; x = np.arange(120 * 5000, dtype=np.uint32).astype(np.uint8); "
            "assert (jp.inverse(jp.forward(x)) == x).all()")
    r = subprocess.run([os.sys.executable, "-c", code], cwd=ROOT, env=dict(os.environ, JP_BWT_TRACE="2"),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "[jp_bwt trace] forward: calls=1" in r.stderr and "[jp_bwt trace] inverse: calls=1" in r.stderr
    assert r.stderr.count("[jp_bwt call]") == 2


def test_fuzz_small_blocks_against_oracle(jp, orc):
    """Randomised differential test: many short blocks of random length, alphabet and structure (runs, repeats,
    zero tails) through forward and inverse, bit-for-bit against the C restatement."""
    rng = np.random.default_rng(20240607)
    for case in range(300):
        n = int(rng.integers(1, 6000)) if case % 7 else int(rng.integers(1, 5)) * 120
        sigma = int(rng.choice([1, 2, 3, 4, 16, 64, 255, 256]))
        T = rng.integers(0, sigma, n).astype(np.uint8)
        kind = case % 5
        if kind == 1 and n > 10:                       # long repeat
            half = n // 2
            T[half: 2 * half] = T[:half]
        elif kind == 2:                                # runs
            T = np.repeat(T[: max(1, n // 17)], 17)[:n]
            n = T.size
        elif kind == 3 and n > 40:                     # zero tail / zero island (sentinel vs 0x00)
            T[-int(rng.integers(1, 40)):] = 0
            T[n // 3: n // 3 + 9] = 0
        elif kind == 4 and n > 300:                    # periodic with a defect
            p = int(rng.integers(1, 40))
            T = np.tile(T[:p], n // p + 1)[:n].copy()
            T[n // 2] ^= 1
        want = orc.forward(T, "port", prefill=0xEE)
        got = jp.forward(T, prefill=0xEE)
        assert (got == want).all(), (case, n, sigma, kind)
        assert (jp.inverse(want) == T).all(), (case, n, sigma, kind)


def test_maximum_block_size_round_trip(jp, orc):
    """format.hpp:22 MAX_BLOCKSIZE = 1000 MiB: 32-bit indices, 31-bit ranks, byte offsets beyond 4 GiB in every table.
    Too large for the CPU oracle to finish in seconds, so the size-independent properties are checked instead."""
    n = 1000 * MiB
    T = orc.gen("markov2", n, 31337)
    out = jp.forward(T)
    st = jp.last_stats()
    nlen = n - n % 120
    assert st.nlen == nlen and st.rounds >= 1
    assert (np.bincount(out[:nlen], minlength=256) == np.bincount(T[:nlen], minlength=256)).all()
    assert (out[nlen:n] == T[nlen:]).all()
    I = orc.indices(out)
    assert I.min() >= 1 and I.max() <= nlen and np.unique(I).size == 120
    back = jp.inverse(out)
    assert back.size == n and (back == T).all()
    assert jp.last_stats().device_bytes <= 4 * nlen + nlen // 32 + (1 << 20)


def test_memory_pressure_degrades_concurrency_not_correctness(jp, orc):
    """Workspaces are per context (forward ~45N). When several large blocks in flight do not fit the device (or
    JP_BWT_DEVICE_MEM_LIMIT), callers take idle contexts' workspaces or wait for a busy one -- no block fails; a
    block that cannot fit at all returns JP_ERR_OOM cleanly."""
    code = r'''
import sys, threading, numpy as np
sys.path.insert(0, %r)
import jampack_b200 as jp, synth
blocks = [synth.gen("markov2", (16 << 20) + 120 * i, 60 + i) for i in range(6)]
outs, errs = [None] * 6, []
def work(i):
    try:
        f = jp.forward(blocks[i]); outs[i] = (synth.fnv(f), bool((jp.inverse(f) == blocks[i]).all()))
    except Exception as e: errs.append(repr(e))
th = [threading.Thread(target=work, args=(i,)) for i in range(6)]
[t.start() for t in th]; [t.join() for t in th]
print("ERRS", errs); print("OUTS", outs)
''' % ROOT
    env = dict(os.environ, JP_BWT_DEVICE_MEM_LIMIT=str(2_000_000_000))
    r = subprocess.run([os.sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert "ERRS []" in r.stdout, r.stdout
    want = [(orc.fnv(orc.forward(orc.gen("markov2", (16 << 20) + 120 * i, 60 + i), _impl(orc))), True) for i in range(6)]
    assert ("OUTS " + repr(want)) in r.stdout, r.stdout
    code2 = ("import numpy as np, jampack_b200 as jp\n"
             "try:\n    jp.forward(np.zeros(16 << 20, dtype=np.uint8)); print('NOERR')\n"
             "except jp.BwtError as e:\n    print('RC', e.rc)\n")
    r = subprocess.run([os.sys.executable, "-c", code2], cwd=ROOT, env=dict(os.environ, JP_BWT_DEVICE_MEM_LIMIT=str(300_000_000)),
                       capture_output=True, text=True, timeout=300)
    assert "RC -4" in r.stdout, r.stdout + r.stderr


def test_corrupted_bwt_bytes_never_hang_or_crash(jp, orc):
    """Flipping bytes in the BWT body breaks the single LF cycle into arbitrary cycles. The decoder must come back
    quickly with an error (or, when the damage happens to leave a consistent permutation, with some block of the
    right size) -- never hang, never read out of bounds; the container's checksum (jampack.cpp:56-57) does the rest."""
    import time
    rng = np.random.default_rng(99)
    T = orc.gen("markov2", MiB, 13)
    B = orc.forward(T, _impl(orc))
    n = T.size
    outcomes = {"error": 0, "output": 0}
    t0 = time.time()
    for trial in range(24):
        bad = B.copy()
        k = int(rng.integers(1, 50)) if trial % 3 else 1
        pos = rng.integers(0, n - n % 120, k)
        bad[pos] = rng.integers(0, 256, k).astype(np.uint8)
        if trial % 4 == 3:
            bad[: n // 2] = bad[n // 2: n // 2 * 2]            # wholesale damage
        try:
            out = jp.inverse(bad)
            assert out.size == n
            outcomes["output"] += 1
        except jp.BwtError as e:
            assert e.rc in (-5, -6), e
            outcomes["error"] += 1
    assert time.time() - t0 < 60
    assert outcomes["error"] > 0
    assert (jp.inverse(B) == T).all()


@pytest.fixture
def single_walk_env():
    """JP_BWT_INV_* are read per call; restore them whatever the test does."""
    keys = ("JP_BWT_INV_SINGLE", "JP_BWT_INV_STREAM_CAP", "JP_BWT_INV_WBLOCKS_PER_SM", "JP_BWT_INV_FLAGS", "JP_BWT_INV_LOG2M")
    saved = {k: os.environ.get(k) for k in keys}
    yield os.environ
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.mark.parametrize("kind,n,seed", [("markov2", 65536 + 120, 3), ("uniform", MiB, 2), ("repetitive", 2 * MiB + 5, 3),
                                         ("alla", MiB + 120, 0), ("markov2", 5 * MiB + 77, 4), ("kat_quadratic", 300000, 0)])
def test_single_walk_inverse_matches_two_pass(jp, orc, single_walk_env, kind, n, seed):
    """The single-walk inverse (decode once into per-warp streams, rank, replay) is the default from 30 Mi up; forced
    on here for small blocks. Same bytes as the two-pass path and as the text, through every entry point."""
    import torch
    T = orc.gen(kind, n, seed)
    B = orc.forward(T, _impl(orc))
    single_walk_env["JP_BWT_INV_SINGLE"] = "0"
    two_pass = jp.inverse(B)
    assert jp.last_stats().stream_chunks == 0
    # the measured-and-rejected variants that are still switches (cache hints of the two-pass walkers, marker spacing)
    for flags, log2m in (("1", None), ("2", None), ("3", "5"), (None, "3")):
        if flags is not None:
            single_walk_env["JP_BWT_INV_FLAGS"] = flags
        if log2m is not None:
            single_walk_env["JP_BWT_INV_LOG2M"] = log2m
        assert (jp.inverse(B) == T).all(), (flags, log2m)
        single_walk_env.pop("JP_BWT_INV_FLAGS", None); single_walk_env.pop("JP_BWT_INV_LOG2M", None)
    single_walk_env["JP_BWT_INV_SINGLE"] = "1"
    one = jp.inverse(B)
    st = jp.last_stats()
    assert st.stream_chunks > 0, "single-walk path did not run"
    assert (one == T).all() and (two_pass == T).all()
    d_in = torch.from_numpy(B.copy()).cuda()
    for consume in (False, True):
        out = jp.inverse_device(d_in.clone(), consume=consume)
        assert jp.last_stats().stream_chunks > 0
        assert (out.cpu().numpy()[:n] == T).all()
    assert (d_in.cpu().numpy() == B).all()                  # the non-consuming entry point left its input alone


def test_single_walk_overflow_reruns_two_pass(jp, orc, single_walk_env):
    """When the stream space runs out the call reruns the two-pass path on the intact LF table: still bit-exact,
    reported as a negative chunk count."""
    T = orc.gen("markov2", 3 * MiB, 17)
    B = orc.forward(T, _impl(orc))
    single_walk_env["JP_BWT_INV_SINGLE"] = "1"
    single_walk_env["JP_BWT_INV_STREAM_CAP"] = "100"
    out = jp.inverse(B)
    assert jp.last_stats().stream_chunks < 0
    assert (out == T).all()
    single_walk_env["JP_BWT_INV_STREAM_CAP"] = "0"
    assert (jp.inverse(B) == T).all() and jp.last_stats().stream_chunks < 0


def test_single_walk_rejects_corrupt_input(jp, orc, single_walk_env):
    """Same error behaviour as the two-pass path: bad indices -5, inconsistent chains -6, never a hang."""
    import time
    single_walk_env["JP_BWT_INV_SINGLE"] = "1"
    rng = np.random.default_rng(5)
    T = orc.gen("markov2", MiB, 21)
    n = T.size
    B = orc.forward(T, _impl(orc))
    bad = B.copy()
    bad[n + 4 * 50: n + 4 * 50 + 4] = np.frombuffer(np.int32(12345).tobytes(), dtype=np.uint8)
    with pytest.raises(jp.BwtError) as e:
        jp.inverse(bad)
    assert e.value.rc == -6
    t0 = time.time()
    errors = 0
    for trial in range(16):
        bad = B.copy()
        pos = rng.integers(0, n - n % 120, int(rng.integers(1, 40)))
        bad[pos] = rng.integers(0, 256, pos.size).astype(np.uint8)
        if trial % 4 == 3:
            bad[: n // 2] = bad[n // 2: n // 2 * 2]
        try:
            assert jp.inverse(bad).size == n
        except jp.BwtError as e2:
            assert e2.rc in (-5, -6), e2
            errors += 1
    assert time.time() - t0 < 60 and errors > 0
    assert (jp.inverse(B) == T).all() and jp.last_stats().stream_chunks > 0


@pytest.mark.parametrize("kind,n,seed,log2", [("markov2", 5 * MiB + 13, 1, 20), ("uniform", 3 * MiB, 2, 19), ("alla", 2 * MiB, 0, 18)])
def test_forward_emits_by_text_region(jp, orc, kind, n, seed, log2):
    """k_fwd_emit_regions (blocks beyond the L2: one sweep over SA per region of the text), here with 256 KiB - 1 MiB regions;
    the 256 MiB golden block takes the same path with its real 64 MiB regions."""
    T = orc.gen(kind, n, seed)
    want = orc.forward(T, _impl(orc), prefill=0x5C)
    saved = os.environ.get("JP_BWT_FWD_EMIT_REGION_LOG2")
    os.environ["JP_BWT_FWD_EMIT_REGION_LOG2"] = str(log2)
    try:
        got = jp.forward(T, prefill=0x5C)
    finally:
        if saved is None:
            os.environ.pop("JP_BWT_FWD_EMIT_REGION_LOG2", None)
        else:
            os.environ["JP_BWT_FWD_EMIT_REGION_LOG2"] = saved
    assert (got == want).all()


@pytest.mark.parametrize("plan", ["i0", "s0", "s4,i0", "i4,s0", "s1,i1,s0"])
@pytest.mark.parametrize("kind,n,seed", [("markov2", 33 * MiB, 1), ("alla", 32 * MiB + 120, 0), ("zeros_in_text", 40 * MiB, 3)])
def test_single_walk_ranking_in_any_block_order(jp, orc, plan, kind, n, seed):
    """The ranking of the single-walk inverse is launched as scattered blocks with a hop budget, then index order to the end
    (bwt_inverse.cu, k_inv_rank_packed); any sequence of orders and budgets must give the same text."""
    T = orc.gen("markov2" if kind == "zeros_in_text" else kind, n, seed)
    if kind == "zeros_in_text":
        T[5 * MiB: 9 * MiB] = 0; T[20 * MiB: 20 * MiB + 70000] = 0
    B = jp.forward(T)
    saved = os.environ.get("JP_BWT_INV_RANK_PLAN")
    os.environ["JP_BWT_INV_RANK_PLAN"] = plan
    try:
        back = jp.inverse(B)
        st = jp.last_stats()
    finally:
        if saved is None:
            os.environ.pop("JP_BWT_INV_RANK_PLAN", None)
        else:
            os.environ["JP_BWT_INV_RANK_PLAN"] = saved
    assert (back == T).all() and st.stream_chunks > 0


def _word_text(n, seed):
    rng = np.random.default_rng(seed)
    words = [bytes(rng.integers(97, 123, int(rng.integers(2, 10)), dtype=np.uint8)) for _ in range(2000)]
    pick = rng.zipf(1.2, n // 3) % len(words)
    return np.frombuffer(b" ".join(words[i] for i in pick), dtype=np.uint8)[:n].copy()


@pytest.mark.parametrize("kind,n,seed", [("markov2", 2 * MiB + 77, 1), ("uniform", MiB, 2), ("words", 3 * MiB, 3), ("dna", 2 * MiB + 1, 4),
                                         ("words+bytes", 4 * MiB, 5), ("markov2", 32 * MiB, 6), ("small", 5000, 7)])
@pytest.mark.parametrize("passes", [None, "4", "6", "8"])
def test_forward_context_coded_keys_are_bit_exact(jp, orc, kind, n, seed, passes):
    """Initial keys as bit strings of context-chosen alphabetic codewords (bwt_forward.cu 2a): forced on, with the
    key length the host picks and with 32-, 48- and 63-bit keys, against the compiled reference; the mixed-radix keys
    (switch off) must give the same block."""
    rng = np.random.default_rng(seed)
    if kind == "words":
        T = _word_text(n, seed)
    elif kind == "dna":
        T = (rng.integers(0, 4, n) + 65).astype(np.uint8)
    elif kind == "words+bytes":
        T = _word_text(n, seed); T[n // 3: n // 3 + 200000] = rng.integers(0, 256, 200000).astype(np.uint8)
    elif kind == "small":
        T = (rng.integers(0, 7, n) + 48).astype(np.uint8)
    else:
        T = orc.gen(kind, n, seed)
    want = orc.forward(T, _impl(orc), prefill=0x5C)
    keys = ("JP_BWT_FWD_CTXKEYS", "JP_BWT_FWD_KEYPASSES")
    saved = {k: os.environ.get(k) for k in keys}
    try:
        os.environ["JP_BWT_FWD_CTXKEYS"] = "1"
        if passes is not None:
            os.environ["JP_BWT_FWD_KEYPASSES"] = passes
        got = jp.forward(T, prefill=0x5C)
        st = jp.last_stats()
        os.environ["JP_BWT_FWD_CTXKEYS"] = "0"
        plain = jp.forward(T, prefill=0x5C)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    assert (got == want).all() and (plain == want).all()
    if kind == "markov2" and n >= 32 * MiB and passes in (None, "6", "8"):
        assert st.active_fraction[0] < 0.1, st.active_fraction[:3]      # the coded keys leave the rounds a few per cent of the block


@pytest.mark.parametrize("kind,n,seed", [("alla", MiB, 0), ("alla", 360, 0), ("repetitive", MiB, 3), ("markov2", MiB + 77, 1),
                                         ("zero_pages", 3 * MiB, 5), ("runs", 2 * MiB, 6), ("tar_like", 4 * MiB, 7), ("alla", 40 * MiB, 0)])
@pytest.mark.parametrize("force", [None, "1"])
def test_forward_run_bypass_is_bit_exact(jp, orc, kind, n, seed, force):
    """Suffixes inside single-symbol runs are placed by counting instead of being sorted (bwt_forward.cu, "run bypass"):
    with the screen deciding (default) and forced on, against the compiled reference."""
    rng = np.random.default_rng(seed)
    if kind == "zero_pages":
        T = rng.integers(0, 256, n).astype(np.uint8)
        for _ in range(12):
            a = int(rng.integers(0, n)); T[a:a + int(rng.integers(1, 1 << 17))] = 0
    elif kind == "runs":
        T = np.ascontiguousarray(np.repeat(rng.integers(0, 5, n // 20 + 1).astype(np.uint8), rng.integers(1, 200, n // 20 + 1))[:n])
    elif kind == "tar_like":           # text files padded with zeros to 512-byte records: many runs of equal length
        T = orc.gen("markov2", n, seed)
        for a in range(0, n - 512, 512 * 7):
            T[a + 300 + (a // 512) % 150: a + 512] = 0
    else:
        T = orc.gen(kind, n, seed)
    want = orc.forward(T, _impl(orc), prefill=0x5C)
    saved = os.environ.get("JP_BWT_FWD_BYPASS")
    if force is not None:
        os.environ["JP_BWT_FWD_BYPASS"] = force
    try:
        got = jp.forward(T, prefill=0x5C)
        st = jp.last_stats()
    finally:
        if saved is None:
            os.environ.pop("JP_BWT_FWD_BYPASS", None)
        else:
            os.environ["JP_BWT_FWD_BYPASS"] = saved
    assert (got == want).all()
    if kind == "alla" and n >= MiB:
        assert st.rounds == 0 and st.bypass_runs == 1 and st.bypass_suffixes > n - 200
        assert (jp.inverse(got) == T).all()


@pytest.mark.parametrize("p,n,seed", [(300, 3 * MiB, 2), (5000, 4 * MiB + 3, 4), (40000, 6 * MiB, 5), (70000, 6 * MiB, 7), (1021, 8 * MiB, 6)])
@pytest.mark.parametrize("packed", ["0", "1"])
def test_forward_periodic_repeats_after_coded_keys(jp, orc, p, n, seed, packed):
    """A periodic block that does NOT go through its representatives (switched off here; periods beyond the probe's 16 Ki never
    do): context-coded initial keys, then the detection after the initial step and the repeat-length keys on groups that
    are finer than "equal h-prefix" (DESIGN 5.4)."""
    rng = np.random.default_rng(seed)
    T = np.tile(rng.integers(0, 5, p).astype(np.uint8) + 97, n // p + 1)[:n].copy()
    T[rng.integers(0, n, 60)] ^= 1
    want = orc.forward(T, _impl(orc), prefill=0x5C)
    keys = ("JP_BWT_FWD_REDUCED", "JP_BWT_FWD_CTXKEYS", "JP_BWT_FWD_PACKED")
    saved = {k: os.environ.get(k) for k in keys}
    try:
        os.environ.update({"JP_BWT_FWD_REDUCED": "0", "JP_BWT_FWD_CTXKEYS": "1", "JP_BWT_FWD_PACKED": packed})
        got = jp.forward(T, prefill=0x5C)
        st = jp.last_stats()
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    assert (got == want).all()
    assert st.period == (p if p <= 65536 else 0), (st.period, st.rounds)     # (the detection looks for distances up to 64 Ki)


@pytest.mark.parametrize("kind,n,seed", [("repetitive", 4 * MiB, 3), ("repetitive", 24 * MiB + 5, 9), ("period_12", 3 * MiB, 2), ("period_5000", 6 * MiB, 4),
                                         ("period_2_clean", 2 * MiB, 0), ("runs_in_a_period", 2 * MiB, 1), ("markov2", 2 * MiB, 1)])
def test_forward_periodic_repeats_are_bit_exact(jp, orc, kind, n, seed):
    """Blocks of period p with sparse defects: the repeat-length keys (bwt_forward.cu, "periodic repeats") against the
    compiled reference, with the detection deciding on its own."""
    rng = np.random.default_rng(seed)
    if kind.startswith("period_"):
        p = int(kind.split("_")[1])
        T = np.tile(rng.integers(0, 4, p).astype(np.uint8) + 97, n // p + 1)[:n].copy()
        if not kind.endswith("clean"):
            T[rng.integers(0, n, 40)] ^= 1
    elif kind == "runs_in_a_period":
        T = np.tile(np.concatenate([np.zeros(700, np.uint8), np.array([1, 2, 1], np.uint8)]), n // 703 + 1)[:n].copy()
        T[rng.integers(0, n, 9)] = 5
    else:
        T = orc.gen(kind, n, seed)
    want = orc.forward(T, _impl(orc), prefill=0x5C)
    got = jp.forward(T, prefill=0x5C)
    st = jp.last_stats()
    assert (got == want).all()
    if kind == "repetitive":
        assert st.period == 1021 and st.rounds <= 10, (st.period, st.rounds)
    if kind == "markov2":
        assert st.period == 0
