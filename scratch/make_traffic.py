#!/usr/bin/env python
"""profiles/traffic.json from one `ncu --set full` capture of the bench command: per bench phase the DRAM bytes (read + write) of its
launches in ONE step, the issue-slot utilisation and the capture's name.  bench.py reads the file for `roofline.traffic`.
usage: python scratch/make_traffic.py gpurun_out/r02j_prof.ncu-rep r02j"""
import csv, json, os, subprocess, sys
rep, tag = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
def col(n): return h.index(n)
def num(r, n):
    v = float(r[col(n)].replace(",", ""))
    u = units[col(n)]
    return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9}.get(u, 1.0)
phase_of = [("k_me_group", "me_search_group"), ("k_me_frac", "me_frac"), ("k_rdoq", "rdoq"), ("k_mc_batch", "mc"), ("k_fwd_tq", "fwd_tq"), ("k_inv_tq", "inv_tq")]
bound = {"me_search_group": "issue slots: per-lane SAD from the staged window (LDS + funnel shift + VABSDIFF4) and the TZ control flow; no table traffic",
         "me_frac": "integer pipe (dp2a vertical taps, Hadamard butterflies)", "rdoq": "issue slots (dependent FP64 chain per TU)",
         "mc": "latency (small launch)", "fwd_tq": "latency / integer pipe (matrix-form transform)", "inv_tq": "latency (global loads, low occupancy at 16x16 / 32x32)"}
acc = {}
for r in rows[2:]:
    name = r[col("Kernel Name")]
    ph = next((p for k, p in phase_of if k in name), None)
    if ph is None:
        continue
    short = name.split("(")[0]
    a = acc.setdefault(ph, {"seen": set(), "bytes": 0.0, "sec": 0.0, "issue": [], "launches": []})
    if short in a["seen"]:
        continue                      # one launch of each distinct kernel = one step's worth (the capture spans more than one step)
    a["seen"].add(short)
    a["bytes"] += num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")
    a["sec"] += num(r, "gpu__time_duration.sum")
    a["issue"].append(round(float(r[col("sm__throughput.avg.pct_of_peak_sustained_elapsed")]), 1))
    a["launches"].append(short)
res = {}
for ph, a in acc.items():
    res[ph] = {"bound": bound[ph], "capture": tag, "dram_bytes_per_launch": a["bytes"], "launches_in_capture": len(a["launches"]),
               "kernels": a["launches"], "ncu_ms": round(a["sec"] * 1e3, 4), "issue_slots_busy_pct": a["issue"],
               "note": "sum over the phase's launches of one step (ncu --set full, cold caches, clocks not locked)"}
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json"), "w"), indent=1, sort_keys=True)
print(json.dumps({k: (v["dram_bytes_per_launch"], v["ncu_ms"], v["issue_slots_busy_pct"]) for k, v in res.items()}, indent=1))
