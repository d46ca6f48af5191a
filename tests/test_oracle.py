"""CPU tests: the oracle restatement (oracle/bwt_oracle.c) against the golden vectors produced by the
unmodified reference (tests/golden/kat.json), against the compiled reference itself when oracle/_ref is
present, and against a brute-force suffix sort on tiny inputs."""
import numpy as np
import pytest

MiB = 1 << 20


def _case_ids(cases):
    return [c["name"] for c in cases]


def test_golden_small_full_vectors(orc, golden):
    for c in golden["cases"]:
        T = orc.gen(c["kind"], c["len"], c["seed"])
        assert "%016x" % orc.fnv(T) == c["fnv_in"], c["name"]
        if "out_hex" not in c:
            continue
        want = np.frombuffer(bytes.fromhex(c["out_hex"]), dtype=np.uint8)
        got = orc.forward(T, "port")
        n, nlen = c["len"], c["nlen"]
        assert (got[: want.size] == want).all(), c["name"]
        if nlen:
            assert [int(x) for x in orc.indices(got)] == c["indices"], c["name"]
            assert (orc.inverse(got, "port") == T).all(), c["name"]
        else:
            assert (got[:n] == T).all()       # len < 120: pure copy, trailer untouched (bwt.cpp:35)


def test_golden_medium_hashes(orc, golden):
    for c in golden["cases"]:
        if "out_hex" in c or c["len"] > 2 * MiB:
            continue
        T = orc.gen(c["kind"], c["len"], c["seed"])
        got = orc.forward(T, "port")
        assert "%016x" % orc.fnv(got[: c["len"]]) == c["fnv_bwt"], c["name"]
        assert "%016x" % orc.fnv(got) == c["fnv_all"], c["name"]
        assert [int(x) for x in orc.indices(got)] == c["indices"], c["name"]
        assert (orc.inverse(got, "port") == T).all()


def test_survey_known_answers(orc):
    """SURVEY.md Appendix B values that do not depend on a hash definition."""
    T = orc.gen("kat_quadratic", 240)
    out = orc.forward(T, "port")
    assert out[:48].tobytes().hex() == ("eeec1a009876d63c04b862ee30ae740aba70c866ac86340212de60e45efa4ce06af8561c"
                                        "f264c298ce90540eea7c501a")
    I = orc.indices(out)
    assert (I[0], I[1], I[119]) == (1, 9, 13)
    T = orc.gen("kat_extremes", 240)
    out = orc.forward(T, "port")
    assert out[:48].tobytes().hex() == ("01ffffffffffffffff0101010101010101010101010101010100000000000000000000000000"
                                        "000000ffffffffffffff")
    I = orc.indices(out)
    assert (I[0], I[1], I[119]) == (81, 24, 65)
    # all-equal input: I[k] = n - k*step (Appendix A edge cases)
    out = orc.forward(orc.gen("alla", 360), "port")
    assert [int(x) for x in orc.indices(out)] == [360 - 3 * k for k in range(120)]
    I = orc.indices(orc.forward(orc.gen("markov2", MiB, 1), "port"))
    assert (I[0], I[1], I[119]) == (829955, 365342, 452592)


def test_short_block_leaves_trailer_untouched(orc):
    T = orc.gen("uniform", 119, 3)
    out = orc.forward(T, "port", prefill=0xAB)
    assert (out[:119] == T).all() and (out[119:] == 0xAB).all()
    back = orc.inverse(out, "port")
    assert (back == T).all()
    out0 = orc.forward(orc.gen("uniform", 0, 0), "port", prefill=0x5A)
    assert out0.size == 480 and (out0 == 0x5A).all()


@pytest.mark.parametrize("n", [120, 121, 239, 240, 1000, 4096])
def test_suffix_array_brute_force(orc, n):
    rng = np.random.default_rng(n)
    for alphabet in (2, 4, 256):
        T = rng.integers(0, alphabet, n).astype(np.uint8)
        if alphabet == 4:
            T[n // 2:] = T[: n - n // 2]           # long repeats
        sa = orc.suffix_array(T)
        b = T.tobytes()
        want = sorted(range(n), key=lambda i: b[i:])
        assert sa.tolist() == want
        assert orc.port().jpo_check_suffix_array(T.ctypes.data_as(orc._u8p), sa.ctypes.data_as(orc._i32p), n) == 0


def test_inverse_is_unit_count_independent(orc):
    T = orc.gen("markov2", 120 * 500 + 77, 11)
    B = orc.forward(T, "port")
    for units in (1, 2, 3, 4, 5, 6, 8, 10, 12, 15, 20, 24, 30, 40, 60, 120):
        assert (orc.inverse(B, "port", units=units) == T).all(), units


def test_inverse_rejects_bad_index(orc):
    T = orc.gen("markov2", 2400, 5)
    B = orc.forward(T, "port").copy()
    B[2400 + 4 * 7: 2400 + 4 * 7 + 4] = np.frombuffer(np.int32(2401).tobytes(), dtype=np.uint8)
    with pytest.raises(RuntimeError):
        orc.inverse(B, "port")


def test_map_is_the_reference_table(orc):
    """Map (bwt.cpp:171-174) drives the reference walk; check its defining properties."""
    T = orc.gen("markov2", 120 * 40, 2)
    B = orc.forward(T, "port")
    nlen = T.size
    idx = int(orc.indices(B)[0])
    Map, Ct = orc.build_map(B[:nlen], nlen, idx)
    rows = np.arange(nlen) + (np.arange(nlen) >= idx)
    assert sorted(Map.tolist()) == rows.tolist()                    # a permutation of the byte rows
    sym = B[:nlen][Map - (Map >= idx)]                              # L symbol of each target row ...
    assert (np.diff(sym.astype(int)) >= 0).all()                    # ... ascending = the F column
    assert Ct[0] == 0 and Ct[256] == nlen and (np.diff(Ct) == np.bincount(B[:nlen], minlength=256)).all()


# ---- the restatement against the compiled reference itself (authoring container / GPU box) -------------
def _need_ref(orc):
    if orc.ref() is None:
        pytest.skip("oracle/_ref not built (needs /root/reference): golden vectors still pin the oracle")


@pytest.mark.parametrize("kind,n,seed", [("uniform", 100003, 1), ("markov2", 300007, 2), ("repetitive", 200000, 3),
                                         ("alla", 50000, 0), ("kat_extremes", 99999, 0), ("markov2", 120, 4),
                                         ("uniform", 119, 5), ("uniform", 1, 6), ("markov2", 2 * MiB + 5, 7)])
def test_port_equals_reference(orc, kind, n, seed):
    _need_ref(orc)
    T = orc.gen(kind, n, seed)
    r = orc.forward(T, "ref", prefill=0x33)
    p = orc.forward(T, "port", prefill=0x33)
    assert (r == p).all()
    assert (orc.inverse(r, "ref", threads=3) == T).all()
    assert (orc.inverse(r, "port") == T).all()


def test_reference_text_with_zero_tail(orc):
    """Blocks ending in runs of 0x00 are where zero-padded keys would tie (SURVEY.md hard parts)."""
    _need_ref(orc)
    T = orc.gen("markov2", 6000, 8)
    T[-300:] = 0
    T[100:200] = 0
    assert (orc.forward(T, "ref") == orc.forward(T, "port")).all()


def test_golden_file_matches_reference(orc, golden):
    _need_ref(orc)
    for c in golden["cases"]:
        if c["len"] > MiB:
            continue
        T = orc.gen(c["kind"], c["len"], c["seed"])
        out = orc.forward(T, "ref")
        assert "%016x" % orc.fnv(out[: c["len"]]) == c["fnv_bwt"], c["name"]
