"""Builds jampack_b200/libjpbwt.so (hand-written sm_100a CUDA + the C-ABI of include/jp_bwt.h) in-tree with nvcc.

    python -m jampack_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libjpbwt.so")
SOURCES = ["jp_bwt_api.cu", "bwt_inverse.cu", "bwt_forward.cu", "src_rle0.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function", "-Xptxas", "-v", "--use_fast_math", "-DRS_SCATTER_MIN_BLOCKS=3"]


def _stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "jp_bwt.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("JP_NVCC_EXTRA", "").split()      # A/B builds: e.g. "-DRS_ITEMS_CFG=12 -DRS_SCATTER_MIN_BLOCKS=4"
    flags = [f for f in NVCC_FLAGS if not any(f.split("=")[0] == e.split("=")[0] for e in extra)] + extra
    env = dict(os.environ)
    env.pop("CC", None), env.pop("CXX", None)   # this image exports a gcc wrapper that nvcc must not pick up
    objs = []
    log = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [nvcc] + flags + ["-ccbin", "/usr/bin/g++", "-I", os.path.join(ROOT, "include"), "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True)
        log.append(r.stderr)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed on " + src)
        objs.append(obj)
    cmd = [nvcc, "-shared", "-ccbin", "/usr/bin/g++", "-o", LIB] + objs + ["-lcudart"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    with open(os.path.join(HERE, "build_ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        sys.stderr.write("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
