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


@pytest.mark.parametrize("kind,n,seed", [("kat_extremes", 2, 0), ("uniform", 3, 1), ("alla", 1000, 0), ("markov2", 3000, 2), ("repetitive", 5000, 3)])
def test_emulated_suffix_array(emu, orc, kind, n, seed):
    """jp::debug_suffix_array (the sorter behind jp_bwt_suffix_array / the -m2 shim) against a brute-force sort."""
    T = orc.gen(kind, n, seed)
    rc, sa = emu.suffix_array(T)
    b = T.tobytes()
    assert rc == 0
    assert sa.tolist() == sorted(range(n), key=lambda i: b[i:])


