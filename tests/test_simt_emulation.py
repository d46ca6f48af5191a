"""CPU tests that run the product's CUDA KERNEL CODE -- jampack_b200/csrc/*.cu, rewritten mechanically by
tests/simt/gen.py (launch syntax, inline PTX) and executed by the SIMT emulator of tests/simt/ -- against the oracle.
This is a checker of the kernels' logic for boxes without a GPU; it is not a CPU path of the product (the C-ABI
library never contains it) and says nothing about speed. The GPU parity tests remain the gate."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import simt  # noqa: E402  (tests/simt: the emulator package)


@pytest.fixture(scope="module")
def emu():
    simt.build()
    return simt


@pytest.fixture
def inv_env():
    keys = ("JP_BWT_INV_SINGLE", "JP_BWT_INV_STREAM_CAP", "JP_BWT_INV_LOG2M")
    saved = {k: os.environ.get(k) for k in keys}
    for k in keys:
        os.environ.pop(k, None)
    yield os.environ
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


INV_CASES = [("kat_quadratic", 240, 0), ("alla", 360, 0), ("kat_extremes", 240, 0), ("uniform", 121, 7), ("markov2", 4093, 9),
             ("repetitive", 70000, 3), ("uniform", 5000, 4), ("markov2", 65536 + 120, 3)]


@pytest.mark.parametrize("kind,n,seed", INV_CASES)
def test_emulated_two_pass_inverse(emu, orc, inv_env, kind, n, seed):
    T = orc.gen(kind, n, seed)
    B = orc.forward(T, "port")
    rc, out, chunks, launches = emu.inverse(B)
    assert rc == 0 and chunks == 0 and launches >= 9
    assert (out == T).all()


@pytest.mark.parametrize("kind,n,seed", [("markov2", 65536 + 120, 3), ("repetitive", 150000, 3), ("alla", 70000, 0), ("uniform", 100000, 2)])
def test_emulated_single_walk_inverse(emu, orc, inv_env, kind, n, seed):
    T = orc.gen(kind, n, seed)
    B = orc.forward(T, "port")
    inv_env["JP_BWT_INV_SINGLE"] = "1"
    for consume in (False, True):
        rc, out, chunks, _ = emu.inverse(B, consume=consume)
        assert rc == 0 and chunks > 0, (rc, chunks)
        assert (out == T).all()
    inv_env["JP_BWT_INV_LOG2M"] = "4"                      # the spacing large blocks use
    rc, out, chunks, _ = emu.inverse(B, consume=True)
    assert rc == 0 and chunks > 0 and (out == T).all()


def test_emulated_single_walk_overflow_reruns_two_pass(emu, orc, inv_env):
    T = orc.gen("markov2", 90000, 5)
    B = orc.forward(T, "port")
    inv_env["JP_BWT_INV_SINGLE"] = "1"
    inv_env["JP_BWT_INV_STREAM_CAP"] = "3"
    rc, out, chunks, launches = emu.inverse(B, consume=True)
    assert rc == 0 and chunks < 0 and launches >= 13
    assert (out == T).all()


@pytest.mark.parametrize("single", ["0", "1"])
def test_emulated_inverse_rejects_corrupt_input(emu, orc, inv_env, single):
    inv_env["JP_BWT_INV_SINGLE"] = single
    T = orc.gen("markov2", 120 * 700, 5)
    n = T.size
    B = orc.forward(T, "port")
    bad = B.copy()
    bad[n + 4 * 7: n + 4 * 7 + 4] = np.frombuffer(np.int32(n + 1).tobytes(), dtype=np.uint8)          # out of range
    assert emu.inverse(bad)[0] == -5
    bad = B.copy()
    bad[n + 4 * 50: n + 4 * 50 + 4] = np.frombuffer(np.int32(12345).tobytes(), dtype=np.uint8)         # wrong but in range
    assert emu.inverse(bad)[0] == -6
    rng = np.random.default_rng(1)
    for _ in range(4):                                                                                  # damaged BWT bytes: an error or some block, never a hang
        bad = B.copy()
        pos = rng.integers(0, n, 20)
        bad[pos] = rng.integers(0, 256, 20).astype(np.uint8)
        assert emu.inverse(bad)[0] in (0, -5, -6)
    assert (emu.inverse(B)[1] == T).all()


FWD_CASES = [("kat_quadratic", 240, 0), ("kat_quadratic", 250, 0), ("alla", 360, 0), ("uniform", 119, 2), ("markov2", 4093, 9),
             ("repetitive", 30000, 3), ("alla", 9000, 0), ("markov2", 65536 + 120, 3)]


@pytest.mark.parametrize("kind,n,seed", FWD_CASES)
def test_emulated_forward(emu, orc, kind, n, seed):
    T = orc.gen(kind, n, seed)
    want = orc.forward(T, "port", prefill=0x5C)
    rc, got, rounds, launches = emu.forward(T)
    assert rc == 0
    assert (got == want).all()
    if n >= 120:
        assert launches > 0


# ---- run bypass: suffixes inside single-symbol runs are placed by counting instead of being sorted ----
def _run_cases():
    rng = np.random.default_rng(37)
    cases = {"all_a": np.full(9000, 97, np.uint8), "tiny_all_a": np.full(360, 5, np.uint8)}
    T = rng.integers(0, 200, 20000).astype(np.uint8); T[5000:15000] = 0
    cases["zero_page_in_noise"] = T
    T = np.zeros(12000, np.uint8); T[4000:] = 1; T[8000:] = 0
    cases["three_plateaus"] = T
    T = rng.integers(0, 4, 8000).astype(np.uint8); T[-500:] = 255; T[:300] = 255
    cases["runs_at_both_ends"] = T
    # two suffixes 1 0^92 2... and 1 0^92 1...: their order hangs on what follows equally long runs
    def seq(*runs):
        return np.concatenate([np.full(l, c, np.uint8) for c, l in runs])
    cases["equal_runs_different_followers"] = np.concatenate([
        seq((2, 40), (1, 1), (0, 92), (2, 10), (1, 89), (0, 80), (2, 57)), seq((1, 33), (2, 5)),
        seq((1, 1), (0, 92), (1, 76), (2, 17), (0, 47), (2, 28)), seq((0, 35), (1, 2), (0, 31), (2, 1))])
    for k in range(4):
        n = 2400 + 700 * k
        cases[f"random_runs_{k}"] = np.ascontiguousarray(np.repeat(rng.integers(0, 2 + k, n // 10 + 1).astype(np.uint8),
                                                                   rng.integers(1, 64 + 10 * k, n // 10 + 1))[:n])
    # 256 symbols (the largest alphabet, base 257 keys), runs of the extreme byte values, equal runs back to back
    T = rng.integers(0, 256, 9000).astype(np.uint8); T[100:400] = 255; T[1000:1300] = 0; T[2000:2300] = 255; T[-64:] = 0
    cases["extreme_bytes_256_symbols"] = T
    cases["many_equal_runs"] = np.tile(np.concatenate([np.zeros(70, np.uint8), np.array([1, 2, 1], np.uint8)]), 60)
    return cases


@pytest.mark.parametrize("name", sorted(_run_cases()))
def test_emulated_forward_with_run_bypass(emu, orc, name):
    T = _run_cases()[name]
    want = orc.forward(T, "port", prefill=0x5C)
    saved = os.environ.get("JP_BWT_FWD_BYPASS")
    try:
        os.environ["JP_BWT_FWD_BYPASS"] = "0"
        rc0, got0, rounds0, _ = emu.forward(T)
        os.environ["JP_BWT_FWD_BYPASS"] = "1"
        rc1, got1, rounds1, _ = emu.forward(T)
    finally:
        if saved is None:
            os.environ.pop("JP_BWT_FWD_BYPASS", None)
        else:
            os.environ["JP_BWT_FWD_BYPASS"] = saved
    assert rc0 == 0 and (got0 == want).all()
    assert rc1 == 0 and (got1 == want).all()
    if name in ("all_a", "tiny_all_a"):
        assert rounds1 == 0 < rounds0, (rounds0, rounds1)      # a block of one repeated byte needs no doubling round at all
    if name in ("zero_page_in_noise", "three_plateaus"):
        assert rounds1 < rounds0, (rounds0, rounds1)


@pytest.mark.parametrize("kind,n,seed", [("markov2", 30000, 1), ("uniform", 9000 + 119, 2), ("repetitive", 20000, 3)])
def test_emulated_forward_emits_by_text_region(emu, orc, kind, n, seed):
    """Blocks beyond the L2 are emitted in one sweep per 64 MiB region of the text (k_fwd_emit_regions); here with 4 KiB regions."""
    T = orc.gen(kind, n, seed)
    want = orc.forward(T, "port", prefill=0x5C)
    saved = os.environ.get("JP_BWT_FWD_EMIT_REGION_LOG2")
    os.environ["JP_BWT_FWD_EMIT_REGION_LOG2"] = "12"
    try:
        rc, got, _, _ = emu.forward(T)
    finally:
        if saved is None:
            os.environ.pop("JP_BWT_FWD_EMIT_REGION_LOG2", None)
        else:
            os.environ["JP_BWT_FWD_EMIT_REGION_LOG2"] = saved
    assert rc == 0 and np.array_equal(got, want)


def _coded_key_blocks():
    """Blocks for the context-coded initial keys: both model orders, a compressible order-1 text (63-bit keys), the end of the
    block inside the last keys, a length that is not a multiple of the key tile."""
    rng = np.random.default_rng(31)
    words = [bytes(rng.integers(97, 123, int(rng.integers(2, 9)), dtype=np.uint8)) for _ in range(300)]
    zipf = rng.zipf(1.3, 40000) % len(words)
    text = np.frombuffer(b" ".join(words[i] for i in zipf), dtype=np.uint8)
    wide = np.concatenate([text[:60000], rng.integers(0, 256, 4000, dtype=np.uint8), text[:30000]])   # > 79 symbols: order 1
    return {
        "order2-dna": rng.integers(0, 4, 40007, dtype=np.uint8) + 65,
        "order2-words": text[:90000 + 13],
        "order1-words-and-bytes": wide,
        "order1-uniform": rng.integers(0, 256, 61000, dtype=np.uint8),
        "short": rng.integers(0, 3, 4200, dtype=np.uint8) + 48,
        "tail-of-equal-symbols": np.concatenate([rng.integers(0, 5, 30000, dtype=np.uint8) + 65, np.full(25, 65, np.uint8)]),
    }


@pytest.mark.parametrize("name", list(_coded_key_blocks()))
@pytest.mark.parametrize("passes", ["", "4", "8", "packed"])
def test_emulated_forward_with_context_coded_keys(emu, orc, name, passes):
    """the initial keys are bit strings of context-chosen alphabetic codewords; the doubling starts from the
    fewest symbols any key covers. Forced on (the default engages it from 1 MiB), with the key length the host picks
    and with the shortest and longest keys; "packed": key and position sorted as one 64-bit word."""
    T = np.ascontiguousarray(_coded_key_blocks()[name])
    want = orc.forward(T, "port", prefill=0x5C)
    keys = ("JP_BWT_FWD_CTXKEYS", "JP_BWT_FWD_KEYPASSES", "JP_BWT_FWD_BYPASS", "JP_BWT_FWD_PACKED")
    saved = {k: os.environ.get(k) for k in keys}
    try:
        os.environ["JP_BWT_FWD_CTXKEYS"] = "1"
        os.environ["JP_BWT_FWD_BYPASS"] = "0"
        os.environ["JP_BWT_FWD_PACKED"] = "1" if passes == "packed" else "0"
        if passes == "packed":
            os.environ.pop("JP_BWT_FWD_KEYPASSES", None)
        elif passes:
            os.environ["JP_BWT_FWD_KEYPASSES"] = passes
        else:
            os.environ.pop("JP_BWT_FWD_KEYPASSES", None)
        rc, got, rounds, launches = emu.forward(T)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    assert rc == 0
    assert np.array_equal(got, want), f"{name}: first difference at {int(np.argmax(got != want))}"


def _periodic(n, p, sig, defects, seed):
    r = np.random.default_rng(seed)
    T = np.tile(r.integers(0, sig, p).astype(np.uint8), n // p + 1)[:n].copy()
    for d in r.integers(0, n, defects):
        T[d] ^= 1
    return T


PERIODIC_CASES = {"repetitive_70k": lambda orc: orc.gen("repetitive", 70000, 3), "period_3": lambda orc: _periodic(12000, 3, 4, 5, 1),
                  "period_7_clean": lambda orc: _periodic(15000, 7, 3, 0, 2), "period_60": lambda orc: _periodic(40000, 60, 5, 10, 3),
                  "period_300_binary": lambda orc: _periodic(90000, 300, 2, 4, 4), "period_2": lambda orc: _periodic(9000, 2, 2, 3, 6),
                  "runs_in_a_period": lambda orc: np.tile(np.concatenate([np.zeros(70, np.uint8), np.array([1, 2, 1], np.uint8)]), 200),
                  "not_periodic": lambda orc: orc.gen("markov2", 30000, 1)}


@pytest.mark.parametrize("name", sorted(PERIODIC_CASES))
def test_emulated_forward_with_periodic_repeats(emu, orc, name):
    """Repeats of period p ordered by repeat length (bwt_forward.cu, "periodic repeats"), alone and on top of the run bypass
    and its jump keys; the detection is forced on (blocks under 64 Ki do not look for a period by default)."""
    T = PERIODIC_CASES[name](orc)
    want = orc.forward(T, "port", prefill=0x5C)
    keys = ("JP_BWT_FWD_BYPASS", "JP_BWT_FWD_PERIODIC", "JP_BWT_FWD_RUNJUMP")
    saved = {k: os.environ.get(k) for k in keys}
    rounds = {}
    try:
        for bp, per in (("0", "0"), ("0", "1"), ("1", "1")):
            os.environ["JP_BWT_FWD_BYPASS"] = bp
            os.environ["JP_BWT_FWD_PERIODIC"] = per
            os.environ["JP_BWT_FWD_RUNJUMP"] = bp
            rc, got, rounds[bp + per], _ = emu.forward(T)
            assert rc == 0 and (got == want).all(), (bp, per)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    if name.startswith(("repetitive", "period_")):
        assert rounds["01"] < rounds["00"], rounds


@pytest.mark.parametrize("name", sorted(PERIODIC_CASES))
def test_emulated_forward_with_periodic_repeats_after_coded_keys(emu, orc, name):
    """The repeat-length keys on top of the context-coded initial keys: the groups they start from are finer than "equal
    h-prefix" and h is the fewest symbols any key covers -- the pass A / pass B argument only needs that much."""
    T = PERIODIC_CASES[name](orc)
    want = orc.forward(T, "port", prefill=0x5C)
    keys = ("JP_BWT_FWD_BYPASS", "JP_BWT_FWD_PERIODIC", "JP_BWT_FWD_REDUCED", "JP_BWT_FWD_CTXKEYS", "JP_BWT_FWD_PACKED")
    saved = {k: os.environ.get(k) for k in keys}
    try:
        os.environ.update({"JP_BWT_FWD_BYPASS": "0", "JP_BWT_FWD_PERIODIC": "1", "JP_BWT_FWD_REDUCED": "0", "JP_BWT_FWD_CTXKEYS": "1"})
        for packed in ("0", "1"):
            os.environ["JP_BWT_FWD_PACKED"] = packed
            rc, got, _, _ = emu.forward(T)
            assert rc == 0 and (got == want).all(), packed
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


REDUCED_CASES = {"repetitive_200k": lambda orc: orc.gen("repetitive", 200000, 3), "period_300_binary": lambda orc: _periodic(160000, 300, 2, 6, 4),
                 "period_60": lambda orc: _periodic(150000, 60, 5, 10, 3), "period_5000": lambda orc: _periodic(180000, 5000, 4, 3, 8),
                 "period_7_clean": lambda orc: _periodic(140000, 7, 3, 0, 2), "not_periodic": lambda orc: orc.gen("markov2", 140000, 1),
                 "half_periodic": lambda orc: np.concatenate([_periodic(100000, 97, 4, 4, 5), orc.gen("markov2", 60000, 2)])}


@pytest.mark.parametrize("name", sorted(REDUCED_CASES))
def test_emulated_forward_of_periodic_block_through_representatives(emu, orc, name):
    """A block whose period shows in the text before the sort is sorted through one representative per stretch and phase
    (bwt_forward.cu, "periodic blocks are sorted through their representatives"); the probe is forced on (blocks under 1 Mi
    do not look by default)."""
    T = REDUCED_CASES[name](orc)
    want = orc.forward(T, "port", prefill=0x5C)
    saved = os.environ.get("JP_BWT_FWD_REDUCED")
    os.environ["JP_BWT_FWD_REDUCED"] = "1"
    try:
        rc, got, rounds, _ = emu.forward(T)
    finally:
        if saved is None:
            os.environ.pop("JP_BWT_FWD_REDUCED", None)
        else:
            os.environ["JP_BWT_FWD_REDUCED"] = saved
    assert rc == 0 and (got == want).all()


@pytest.mark.parametrize("kind,n,seed", [("kat_extremes", 2, 0), ("uniform", 3, 1), ("alla", 1000, 0), ("markov2", 3000, 2), ("repetitive", 5000, 3)])
def test_emulated_suffix_array(emu, orc, kind, n, seed):
    """jp::debug_suffix_array (the sorter behind jp_bwt_suffix_array / the -m2 shim) against a brute-force sort."""
    T = orc.gen(kind, n, seed)
    rc, sa = emu.suffix_array(T)
    b = T.tobytes()
    assert rc == 0
    assert sa.tolist() == sorted(range(n), key=lambda i: b[i:])


