#!/usr/bin/env python
"""tools/measure_int8_peak.py — run tools/micro/umma_peak (148 persistent CTAs issuing tcgen05.mma.kind::i8 N=256) on the
GPU box with an nvidia-smi clock / power record beside it, and write the result as JSON.

    python tools/measure_int8_peak.py [seconds] > gpurun_out/int8_peak.json

The committed copy (profiles/int8_peak_r2.json) is what bench.py reads for `roofline.peak_source = measured_int8`."""
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
    "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"


def main():
    secs = sys.argv[1] if len(sys.argv) > 1 else "2.0"
    exe = os.path.join(ROOT, "tools", "micro", "umma_peak")
    if not os.path.exists(exe):
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3",
                               "-I" + os.path.join(ROOT, "yolo-compression-and-deployment-in-fpga_b200", "csrc"),
                               "-o", exe, exe + ".cu"])
    rows = []
    smi = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=" + Q, "--format=csv,noheader,nounits", "-lms", "50"],
                           stdout=subprocess.PIPE, text=True)

    def rd():
        for line in smi.stdout:
            rows.append((time.time(), [c.strip() for c in line.split(",")]))
    t = threading.Thread(target=rd, daemon=True)
    t.start()
    time.sleep(0.3)
    t0 = time.time()
    out = subprocess.run([exe, secs], stdout=subprocess.PIPE, text=True, check=True).stdout
    t1 = time.time()
    time.sleep(0.2)
    smi.terminate()
    legs = [json.loads(l) for l in out.splitlines() if l.startswith("{")]
    during = [r for (ts, r) in rows if t0 <= ts <= t1 and len(r) >= 7]
    sm = sorted(float(r[0]) for r in during if r[0].replace(".", "").isdigit())
    pw = [float(r[2]) for r in during if r[2].replace(".", "").isdigit()]
    reasons = sorted({n for r in during for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7])
                      if v.lower().startswith("active")})
    best = {}
    for l in legs:
        k = l["leg"]
        if k not in best or l["tops"] > best[k]["tops"]:
            best[k] = l
    res = {
        "what": "tcgen05.mma.kind::i8 M128(xCTAs) N256 K32, one persistent CTA per SM, operands resident in shared memory, accumulators in TMEM; "
                "TOPS = 2*M*N*K*instructions / CUDA-event time (tools/micro/umma_peak.cu)",
        "int8_tops_burst": best.get("burst", {}).get("tops"), "int8_tops_sustained": best.get("sustained", {}).get("tops"),
        "legs": legs,
        "clocks": {"samples": len(sm), "sm_mhz_min": sm[0] if sm else None, "sm_mhz_median": sm[len(sm) // 2] if sm else None,
                   "sm_mhz_max": sm[-1] if sm else None, "power_w_max": max(pw) if pw else None, "reasons": reasons},
        "gpu": subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=name,driver_version", "--format=csv,noheader"], stdout=subprocess.PIPE, text=True).stdout.strip(),
        "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
    }
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
