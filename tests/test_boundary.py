"""CPU tests of the drop-in boundary: the C-ABI library builds, loads, exports every symbol include/jp_bwt.h
declares, and refuses to compute without a GPU (no CPU path)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def jp():
    import jampack_b200
    from jampack_b200 import build
    build.build()
    return jampack_b200


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "jp_bwt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jp_(?:bwt|src)_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_all_exported(jp):
    L = ctypes.CDLL(jp.LIB_PATH)
    names = _declared_symbols()
    assert "jp_bwt_forward" in names and "jp_bwt_inverse" in names and len(names) >= 12
    for n in names:
        assert hasattr(L, n), n
    assert sorted(jp.EXPORTS) == names


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "jp_bwt.h"\nint main(void){ jp_bwt_stats s; (void)s; return JP_BWT_UNITS == 120 ? 0 : 1; }\n')
    import subprocess
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                    "-o", str(tmp_path / "t.o")], check=True)


def test_stats_struct_layout_matches_header(jp, tmp_path):
    src = tmp_path / "s.c"
    src.write_text('#include <stdio.h>\n#include "jp_bwt.h"\nint main(void){ printf("%zu", sizeof(jp_bwt_stats)); return 0; }\n')
    import subprocess
    exe = tmp_path / "s"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    size = int(subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout)
    assert size == ctypes.sizeof(jp.Stats)


def test_no_cpu_fallback(jp):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert jp.device_count() == 0
    with pytest.raises(jp.BwtError) as e:
        jp.forward(np.arange(1000, dtype=np.uint8))
    assert e.value.rc == -2
    with pytest.raises(jp.BwtError):
        jp.inverse(np.zeros(1000, dtype=np.uint8))
    with pytest.raises(jp.BwtError):
        jp.Bwt().ForwardBwt(jp.Buffer(np.zeros(2000, dtype=np.uint8), 1000), jp.Buffer(np.zeros(2000, dtype=np.uint8), 0))


def test_argument_errors(jp):
    L = jp.lib()
    ol = ctypes.c_int32(0)
    assert L.jp_bwt_forward(None, 10, None, ctypes.byref(ol)) == -1
    buf = np.zeros(100, dtype=np.uint8)
    assert L.jp_bwt_inverse(buf.ctypes.data, 100, buf.ctypes.data, ctypes.byref(ol)) == -1   # shorter than a trailer
    assert b"argument" in L.jp_bwt_strerror(-1)
    assert b"no CUDA device" in L.jp_bwt_strerror(-2)


def test_product_does_not_import_oracle():
    """The oracle is a checker only: nothing under jampack_b200/ or include/ may reference it."""
    for base in ("jampack_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                    text = open(os.path.join(dp, f), errors="ignore").read()
                    assert "oracle" not in text.lower() or f == "bwt_shim.cpp", os.path.join(dp, f)


def test_host_shims_compile_standalone(tmp_path):
    """The C++ host side a Jampack maintainer links: both shims build without the reference's headers
    (-DJP_STANDALONE_STAGE uses jampack_b200/host/jp_stage.hpp) and define exactly the reference's symbols."""
    import subprocess
    host = os.path.join(ROOT, "jampack_b200", "host")
    objs = []
    for src, extra in (("bwt_shim.cpp", ["-DJP_STANDALONE_STAGE"]), ("divsufsort_shim.cpp", [])):
        obj = tmp_path / (src + ".o")
        subprocess.run(["g++", "-std=c++14", "-Wall", "-Werror", "-c", os.path.join(host, src), "-I", host,
                        "-I", os.path.join(ROOT, "include"), "-o", str(obj)] + extra, check=True)
        objs.append(str(obj))
    syms = subprocess.run(["nm", "-C", "--defined-only"] + objs, capture_output=True, text=True, check=True).stdout
    assert "BlockSort::Bwt::ForwardBwt(Buffer, Buffer)" in syms
    assert "BlockSort::Bwt::InverseBwt(Buffer, Buffer, Options)" in syms
    assert " T divsufsort" in syms
    und = subprocess.run(["nm", "-u"] + objs, capture_output=True, text=True, check=True).stdout
    for s in ("jp_bwt_forward", "jp_bwt_inverse", "jp_bwt_suffix_array", "jp_bwt_strerror"):
        assert s in und


def test_missing_library_fails_loudly():
    """No silent fallback when the CUDA library has not been built: the first call raises, naming the file."""
    code = ("import jampack_b200 as jp, numpy as np\n"
            "jp.LIB_PATH = '/nonexistent/libjpbwt.so'; jp._lib = None\n"
            "try:\n    jp.forward(np.zeros(1000, dtype=np.uint8))\nexcept jp.BwtError as e:\n    print('RC', e.rc, 'libjpbwt.so' in str(e))\n")
    r = subprocess.run([os.sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert "RC -2 True" in r.stdout, r.stdout + r.stderr
