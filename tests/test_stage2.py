"""Second stage, first half: sorted rank coding + RLE0 per 1 MiB chunk.
CPU: the C restatement against the golden vectors (generated from the compiled reference) and against the reference itself
where oracle/_ref exists; the product's kernel code under the SIMT emulator against the restatement.
GPU: jp_src_rle0 / jp_src_rle0_device (on a block that jp_bwt_forward_device left in HBM) against the reference."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
MiB = 1 << 20


def _golden():
    return json.load(open(os.path.join(HERE, "golden", "stage2.json")))["cases"]


def _same(a, b):
    return len(a) == len(b) and all(x.size == y.size and (x == y).all() for x, y in zip(a, b))


def _check_against_golden(orc, c, freq, rle):
    assert len(rle) == c["chunks"] and [int(r.size) for r in rle] == c["rlen"], c["kind"]
    assert "%016x" % orc.fnv(np.ascontiguousarray(freq).view(np.uint8).ravel()) == c["fnv_freq"], c["kind"]
    assert ["%016x" % orc.fnv(np.ascontiguousarray(r).view(np.uint8)) for r in rle] == c["fnv_rle"], c["kind"]
    if "rle" in c:
        assert [int(x) for x in rle[0]] == c["rle"]
        assert {str(i): int(v) for i, v in enumerate(freq[0]) if v} == c["freq_nonzero"]


def test_restatement_matches_golden_vectors(orc):
    for c in _golden():
        B = orc.forward(orc.gen(c["kind"], c["len"], c["seed"]), "port")
        assert B.size == c["block_bytes"]
        freq, rle = orc.src_rle0(B, "port")
        _check_against_golden(orc, c, freq, rle)


def test_restatement_matches_compiled_reference(orc):
    if orc.ref() is None or not hasattr(orc.ref(), "ref_src_rle0"):
        pytest.skip("oracle/_ref was not built with rank.cpp / rle.cpp")
    rng = np.random.default_rng(3)
    blocks = [orc.forward(orc.gen("markov2", 200000, 5), "port"), rng.integers(0, 256, MiB + 999).astype(np.uint8), np.zeros(5000, np.uint8),
              np.array([7], np.uint8), rng.integers(0, 3, 70000).astype(np.uint8), np.arange(256, dtype=np.uint8).repeat(3)]
    for B in blocks:
        fp, rp = orc.src_rle0(B, "port")
        fr, rr = orc.src_rle0(B, "ref")
        assert (fp == fr).all() and _same(rp, rr)


@pytest.mark.parametrize("kind,n,seed", [("kat_quadratic", 240, 0), ("alla", 360, 0), ("markov2", 4093, 9), ("uniform", 70000, 2), ("markov2", 120000, 1),
                                         ("repetitive", 150000, 3), ("markov2", MiB + 70000, 4)])
def test_emulated_kernels_match_restatement(orc, kind, n, seed):
    import simt
    simt.build()
    B = orc.forward(orc.gen(kind, n, seed), "port")
    rc, freq, rle = simt.src_rle0(B)
    fp, rp = orc.src_rle0(B, "port")
    assert rc == 0 and (freq == fp).all() and _same(rle, rp)


def test_emulated_kernels_on_raw_bytes(orc):
    """Not only BWT output: every byte value, long zero runs, a block that ends one byte into a new chunk."""
    import simt
    simt.build()
    rng = np.random.default_rng(11)
    T = rng.integers(0, 256, 50000).astype(np.uint8); T[1000:30000] = 0; T[-1] = 0
    for B in (T, np.zeros(9000, np.uint8), np.arange(256, dtype=np.uint8).repeat(5), np.concatenate([rng.integers(0, 2, MiB).astype(np.uint8), np.array([9], np.uint8)])):
        rc, freq, rle = simt.src_rle0(B)
        fp, rp = orc.src_rle0(B, "port")
        assert rc == 0 and (freq == fp).all() and _same(rle, rp)


@pytest.mark.gpu
def test_gpu_stage2_matches_golden_and_reference(jp, orc):
    impl = "ref" if (orc.ref() is not None and hasattr(orc.ref(), "ref_src_rle0")) else "port"
    for c in _golden():
        B = jp.forward(orc.gen(c["kind"], c["len"], c["seed"]))
        freq, rle = jp.src_rle0(B)
        _check_against_golden(orc, c, freq, rle)
        fr, rr = orc.src_rle0(B, impl)
        assert (freq == fr).all() and _same(rle, rr)
        assert jp.last_stats().kernel_launches == 4


@pytest.mark.gpu
@pytest.mark.parametrize("kind,mib,seed", [("markov2", 64, 1), ("uniform", 16, 2), ("alla", 8, 0)])
def test_gpu_stage2_consumes_the_bwt_in_hbm(jp, orc, kind, mib, seed):
    """forward_device -> src_rle0_device on the same resident block (the BWT never leaves HBM), against the reference's
    Postcoder::Encode + RLE::encode on the reference's own BWT of the same text."""
    import torch
    impl = "ref" if (orc.ref() is not None and hasattr(orc.ref(), "ref_src_rle0")) else "port"
    T = orc.gen(kind, mib * MiB, seed)
    d_B = jp.forward_device(torch.from_numpy(T).cuda())
    freq, rle, rlen = jp.src_rle0_device(d_B)
    st = jp.last_stats()
    want_f, want_r = orc.src_rle0(orc.forward(T, "ref" if orc.ref() is not None else "port"), impl)
    assert (freq.cpu().numpy() == want_f).all()
    got = rle.cpu().numpy().view(np.uint16)
    rl = rlen.cpu().numpy()
    assert [int(x) for x in rl] == [int(r.size) for r in want_r]
    for k, w in enumerate(want_r):
        assert (got[k * MiB: k * MiB + w.size] == w).all(), (kind, k)
    assert st.direction == 2 and st.ms_total > 0
