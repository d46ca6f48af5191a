"""Prints one line per kernel of jampack_b200/libjpbwt.so: registers, instruction count and a hash of its SASS with the
kernel-parameter offsets and absolute addresses masked. Two builds whose lines agree run the same device code -- the
check used before shipping a switch that is off by default (profiles/sass_r01.txt = the build measured in round 1).
    python tools/sass_fingerprint.py [lib.so] > out.txt ; diff profiles/sass_r01.txt out.txt"""
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "jampack_b200", "libjpbwt.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], capture_output=True, text=True, check=True).stdout
regs = {}
cur = None
for ln in res.splitlines():
    m = re.search(r"Function (\S+):", ln)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+)", ln)
    if m and cur:
        regs[cur] = int(m.group(1))
kern, cur = {}, None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1); kern[cur] = []; continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
        body = re.sub(r"/\*[0-9a-f]+\*/", "", ln).strip()
        body = re.sub(r"c\[0x0\]\[0x[0-9a-f]+\]", "c[P]", body)
        body = re.sub(r"0x[0-9a-f]{6,}", "ADDR", body)
        kern[cur].append(body)
for name in sorted(kern):
    demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
    h = hashlib.sha256("\n".join(kern[name]).encode()).hexdigest()[:16]
    print(f"{demangled:60s} regs={regs.get(name, -1):3d} instr={len(kern[name]):5d} sass={h}")
