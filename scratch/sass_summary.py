#!/usr/bin/env python
"""Per-kernel SASS summary of libthevc_cuda.so (cuobjdump -sass + -res-usage): the mnemonics that prove the B200 features used
(UTMALDG = TMA tile loads, SYNCS = mbarrier, VABSDIFF4 = SIMD absolute differences, IDP = dp2a / dp4a), the widths of the
global / shared accesses and local-memory traffic (spills).  usage: python scratch/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "thevc_b200", "lib", "libthevc_cuda.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
regs = {}
cur = None
for ln in res.splitlines():
    m = re.match(r"\s*Function (\S+):", ln)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", ln)
    if m and cur:
        regs[cur] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
keys = ["UTMALDG", "SYNCS", "VABSDIFF4", "IDP", "LDS.128", "LDS.64", "LDG.E.128", "LDG.E.64", "STG.E.128", "STG.E.64", "ATOMS", "SHFL", "STL", "LDL", "DFMA", "DMUL", "DADD"]
fn = None
cnt = collections.OrderedDict()
arch = set()
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        fn = m.group(1)
        cnt[fn] = collections.Counter()
        continue
    m = re.match(r"\s*arch = (\S+)", ln)
    if m:
        arch.add(m.group(1))
    if fn is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if not m:
        continue
    op = m.group(1)
    cnt[fn]["total"] += 1
    for k in keys:
        if op == k or op.startswith(k + ".") or (k in ("LDS.128", "LDS.64", "LDG.E.128", "LDG.E.64", "STG.E.128", "STG.E.64") and op.startswith(k.split(".")[0]) and k.split(".", 1)[1] in op):
            cnt[fn][k] += 1
print("libthevc_cuda.so: %d kernels, arch %s" % (len(cnt), ",".join(sorted(arch))))
print("%-78s %6s %5s %6s | %s" % ("kernel (demangled prefix)", "instr", "regs", "local", " ".join("%s" % k for k in keys)))
for f, c in sorted(cnt.items(), key=lambda kv: -kv[1]["total"]):
    dem = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip().split("(")[0][:78]
    r = regs.get(f, (0, 0, 0))
    print("%-78s %6d %5d %6d | %s" % (dem, c["total"], r[0], r[2], " ".join("%d" % c[k] for k in keys)))
tot = collections.Counter()
for c in cnt.values():
    tot.update(c)
print("TOTAL: " + " ".join("%s=%d" % (k, tot[k]) for k in keys))
print("tcgen05 / TMEM mnemonics (UTCMMA, UTCHMMA, LDTM, STTM): %d -- nothing on this path is a GEMM (integer / byte work; DESIGN.md 3)" %
      sum(1 for ln in sass.splitlines() if re.search(r"\b(UTC\w*MMA|LDTM|STTM)\b", ln)))
