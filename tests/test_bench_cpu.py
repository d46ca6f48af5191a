"""CPU tests of bench.py's baseline legs: the reference arm's JSON line, the single-block all-core shape and the
legacy-CUDA baseline's behaviour when there is no GPU (it must report, never raise)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MiB = 1 << 20


@pytest.fixture(scope="module")
def bench():
    sys.path.insert(0, ROOT)
    import bench as b
    return b


def test_single_block_shape_matches_reference_output(bench, orc):
    if orc.ref() is None:
        pytest.skip("oracle/_ref not built")
    T = orc.gen("markov2", MiB, 1)
    B = orc.forward(T, "ref")
    inv = bench.cpu_reference_single_block("inverse", T, B, 4)
    fwd = bench.cpu_reference_single_block("forward", T, B, 4)
    assert inv["output_matches"] and fwd["output_matches"]
    assert inv["value"] > 0 and fwd["value"] > 0 and inv["threads"] == 4


def test_legacy_cuda_baseline_reports_instead_of_raising(bench, orc):
    import torch
    T = orc.gen("markov2", 120 * 2000, 2)
    B = orc.forward(T, "ref" if orc.ref() is not None else "port")
    r = bench.legacy_cuda_baseline(T, B)
    assert isinstance(r, dict)
    if torch.cuda.is_available() and "value" in r:
        assert r["output_matches"] and r["value"] > 0
    else:
        assert "unavailable" in r or "value" in r


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--block-mib", "1"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "inv BWT MB/s" and d["unit"] == "MB/s"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
