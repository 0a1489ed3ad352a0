#!/usr/bin/env python
"""Summarise an ncu report's SASS source page for one launch: stall samples per warp role of the tcgen05 kernels
(regions are delimited by marker instructions) and the hottest instructions.  Usage: ncu_regions.py rep.ncu-rep [launch_index] [kernel regex]"""
import csv, subprocess, sys
rep = sys.argv[1]; k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rx = sys.argv[3] if len(sys.argv) > 3 else "conv3x3"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", str(k), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:100])
hdr = rows[1]
data = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[0].startswith('0x')]
seen = set(); uniq = []
for r in data:
    if r[0] in seen: continue
    seen.add(r[0]); uniq.append(r)
data = uniq
iS, iI, iSrc = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Source')
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[iS]) for r in data)
print("sass", len(data), "samples", tot, "warp-inst", sum(int(r[iI]) for r in data))
d = {}
for h in stalls:
    j = hdr.index(h); d[h] = sum(int(r[j] or 0) for r in data)
print({k: v for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:8]})
top = sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:int(sys.argv[4]) if len(sys.argv) > 4 else 25]
for i in sorted(top):
    r = data[i]; print(i, r[iS], r[iI], r[iSrc][:100])
if len(sys.argv) > 5:
    step = int(sys.argv[5])
    for a in range(0, len(data), step):
        b = min(a + step, len(data))
        print(a, "inst", sum(int(r[iI]) for r in data[a:b]), "samples", sum(int(r[iS]) for r in data[a:b]), "|", data[a][iSrc][:50])
