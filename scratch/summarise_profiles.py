#!/usr/bin/env python
"""Turn gpurun_out/ ncu outputs into the tracked summaries under profiles/.

  python scratch/summarise_profiles.py <tag> <launches.csv> <full.ncu-rep> "<command the captures ran>" [more.ncu-rep ...]

writes profiles/<tag>_launches.csv (copy), profiles/<tag>_launches_summary.txt (per-kernel share of the step),
profiles/<tag>_kernels_full.txt (key counters of every kernel in the --set full capture) and updates
profiles/traffic.json (dram bytes per launch per bench phase; bench.py reports it as roofline.traffic)."""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PHASE_OF = {"k_me_sad_tables": "me_tables", "k_me_raster": "me_raster", "k_me_search": "me_search", "k_me_frac": "me_frac",
            "k_rdoq": "rdoq", "k_mc_batch": "mc", "k_fwd_tq": "fwd_tq", "k_inv_tq": "inv_tq"}
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
           "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed.avg.per_cycle_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def short(name):
    n = name.split("(")[0]
    return n.replace("void ", "").replace("tvc::", "").strip()


def main():
    tag, launches, rep, cmd = sys.argv[1:5]
    out = os.path.join(ROOT, "profiles")
    os.makedirs(out, exist_ok=True)
    shutil.copy(launches, os.path.join(out, tag + "_launches.csv"))
    lines = [l for l in open(launches) if not l.startswith("==")]
    r = csv.DictReader(io.StringIO("".join(lines)))
    t = collections.OrderedDict()
    n = 0
    for row in r:
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = short(row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        if row["Metric Unit"] in ("us", "usecond"):
            v *= 1e3
        elif row["Metric Unit"] in ("ms", "msecond"):
            v *= 1e6
        d = t.setdefault(k, [0, 0.0])
        d[0] += 1; d[1] += v; n += 1
    total = sum(v[1] for v in t.values())
    with open(os.path.join(out, tag + "_launches_summary.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none: %s\n" % cmd)
        f.write("%d launches, total %.3f ms (cold-cache, serialised; compare shares)\n" % (n, total / 1e6))
        for k, (c, ns) in sorted(t.items(), key=lambda kv: -kv[1][1]):
            f.write("%-60s launches %4d  time %10.3f ms  share %5.1f%%\n" % (k, c, ns / 1e6, 100 * ns / total))
    # full capture
    hdr, units, rows = None, None, []
    for one in [rep] + sys.argv[5:]:          # further reports (single-kernel captures of other commands) are appended
        q = subprocess.run(["ncu", "-i", one, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)], capture_output=True, text=True)
        rr = list(csv.reader(io.StringIO(q.stdout)))
        if hdr is None:
            hdr = rr[0]
        assert rr[0] == hdr, "reports with different columns"
        rows += [(r_, rr[1]) for r_ in rr[2:]]          # units differ between reports (us / ms, Kbyte / Gbyte)
    ki = hdr.index("Kernel Name")
    traffic_path = os.path.join(out, "traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    per_phase = collections.defaultdict(lambda: [0, 0.0, 0.0])
    issue = collections.defaultdict(list)
    BOUND = {"me_tables": "hbm write", "me_raster": "shared-memory scans", "me_search": "latency (L2/HBM reads, 125 registers -> 16 warps/SM)",
             "me_frac": "integer pipe (issue slots)", "rdoq": "latency (dependent FP64 chain)", "mc": "hbm", "fwd_tq": "hbm", "inv_tq": "hbm"}
    with open(os.path.join(out, tag + "_kernels_full.txt"), "w") as f:
        f.write("ncu --set full --clock-control none --import-source on: %s\n(one launch per kernel; under the profiler -- not a bench value)\n\n" % cmd)
        for row, units in rows:
            name = short(row[ki])
            f.write("%s  grid %s block %s\n" % (row[ki][:110], row[hdr.index("Grid Size")], row[hdr.index("Block Size")]))
            vals = {}
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    vals[m] = (row[i], units[i])
                    f.write("    %-78s %14s %s\n" % (m, row[i], units[i]))
            f.write("\n")
            base = name.split("<")[0]
            ph = PHASE_OF.get(base)
            if ph and "dram__bytes_read.sum" in vals:
                def to_bytes(v, u):
                    x = float(v.replace(",", ""))
                    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                b = to_bytes(*vals["dram__bytes_read.sum"]) + to_bytes(*vals["dram__bytes_write.sum"])
                per_phase[ph][0] += 1; per_phase[ph][1] += b
                if "smsp__issue_active.avg.pct_of_peak_sustained_active" in vals:
                    issue[ph].append(float(vals["smsp__issue_active.avg.pct_of_peak_sustained_active"][0].replace(",", "")))
    for ph, (c, b, _) in per_phase.items():
        e = traffic.get(ph) if isinstance(traffic.get(ph), dict) else {}
        e.update({"dram_bytes_per_launch": b / c, "launches_in_capture": c, "capture": tag, "bound": BOUND.get(ph)})
        if issue[ph]:
            e["issue_slots_busy_pct"] = [round(v, 1) for v in issue[ph]]
        traffic[ph] = e
    json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)
    print("wrote profiles/%s_* and traffic.json" % tag)


if __name__ == "__main__":
    main()
