#!/usr/bin/env python
"""Turn gpurun_out/ captures into the tracked summaries under profiles/ (round given as argv[1], default r1):
launch list csv, per-kernel table of the `ncu --set full` step capture (json + markdown), DRAM traffic json."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r1"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
names = ["conv1", "conv2", "conv3_1", "conv3_2", "conv4_1", "conv4_2", "conv5", "conv6", "conv7", "pred", "head"]     # (the decode is fused into the NMS kernel: 11 launches per step)
raw = subprocess.run(["ncu", "-i", os.path.join(G, "prof_%s_step.ncu-rep" % rnd), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr = rows[0]; units = rows[1]
t_unit = units[hdr.index('gpu__time_duration.sum')]
t_us = {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3, 's': 1e6}.get(t_unit, 1e-3)
b_unit = units[hdr.index('dram__bytes_read.sum')]
b_mb = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(b_unit, 1.0)
keys = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max']
data = rows[2:]
assert len(data) == len(names) + 1, len(data)          # first captured launch = the previous step's NMS
out = []
for n, r in zip(["head (previous step)"] + names, data):
    d = {k: r[hdr.index(k)] for k in keys}; d["layer"] = n; d["units"] = {"time": t_unit, "dram_bytes": b_unit}; out.append(d)
json.dump(out, open(os.path.join(P, "step_%s_ncu_full.json" % rnd), "w"), indent=1)
tr = {"batch": 256, "source": "profiles/step_%s_ncu_full.json (ncu --set full --clock-control none, second step of tools/profile_step.py 2 256)" % rnd,
      "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)"}
md = ["| kernel | layer | time us | DRAM read MB | DRAM write MB | DRAM % | tensor pipe % | issue slots % | regs |", "|---|---|---|---|---|---|---|---|---|"]
for d in out[1:]:
    rd, wr = float(d['dram__bytes_read.sum']) * b_mb, float(d['dram__bytes_write.sum']) * b_mb
    if d["layer"] != "head_decode":
        tr[d["layer"]] = int((rd + wr) * 1e6)
    md.append("| `%s` | %s | %.0f | %.0f | %.0f | %.1f | %.1f | %.0f | %s |" % (
        d['Kernel Name'].split('(')[0].replace('void ', '')[:44], d["layer"], float(d['gpu__time_duration.sum']) * t_us, rd, wr,
        float(d['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']), float(d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']),
        float(d['smsp__issue_active.avg.pct_of_peak_sustained_active']), d['launch__registers_per_thread']))
json.dump(tr, open(os.path.join(P, "traffic_%s.json" % rnd), "w"), indent=1)
open(os.path.join(P, "step_%s_table.md" % rnd), "w").write("\n".join(md) + "\n")
import shutil
shutil.copy(os.path.join(G, "launches.csv"), os.path.join(P, "launches_%s.csv" % rnd))
for src, dst in (("bench.log", "bench_%s_n1.json" % rnd), ("bench_ref.log", "bench_%s_reference.json" % rnd)):
    lines = [l for l in open(os.path.join(G, src)) if l.startswith("{")]
    open(os.path.join(P, dst), "w").write(lines[-1])
for n in (2, 4, 8):
    f = os.path.join(G, "bench_n%d.log" % n)
    if os.path.exists(f):
        lines = [l for l in open(f) if l.startswith("{")]
        if lines:
            open(os.path.join(P, "bench_%s_n%d.json" % (rnd, n)), "w").write(lines[-1])
if os.path.exists(os.path.join(G, "clocks_after_profiles.csv")):
    shutil.copy(os.path.join(G, "clocks_after_profiles.csv"), os.path.join(P, "clocks_%s.csv" % rnd))
print("\n".join(md))
