#!/usr/bin/env python
"""Opcode histogram of the tensor-core / TMA / TMEM instructions per kernel of libyolo_b200.so (cuobjdump -sass):
the SASS evidence that the contraction kernels are tcgen05 (UTCIMMA), TMEM (LDTM), TMA (UTMALDG) and bulk-copy (UBLKCP) code.
usage: python tools/sass_hist.py > profiles/sass_r2.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "yolo-compression-and-deployment-in-fpga_b200", "libyolo_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
OPS = ["UTCIMMA", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "IMMA", "HMMA", "IDP", "LDGSTS", "FFMA2", "FADD2", "I2IP", "REDUX", "SHFL"]
per, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        per[cur]["total"] += 1
        for o in OPS:
            if op == o or op.startswith(o + "."):
                per[cur][o] += 1
def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    except Exception:
        return n
tot = collections.Counter()
print("SASS opcode histogram of %s (sm_100a), one row per kernel; columns with a zero total are omitted" % os.path.basename(so))
cols = [o for o in OPS if any(c[o] for c in per.values())]
print("%-110s %7s " % ("kernel", "instrs") + " ".join("%8s" % o for o in cols))
for n, c in per.items():
    if not any(c[o] for o in ("UTCIMMA", "IMMA", "IDP", "UTMALDG", "UBLKCP", "LDTM")):
        continue
    d = re.sub(r"^void ", "", demangle(n))
    d = re.sub(r"\(.*\)$", "", d)
    print("%-110s %7d " % (d[:110], c["total"]) + " ".join("%8d" % c[o] for o in cols))
    tot.update(c)
print("%-110s %7d " % ("TOTAL (kernels listed)", tot["total"]) + " ".join("%8d" % tot[o] for o in cols))
