#!/usr/bin/env python
"""Per-CUDA-source-line totals of one kernel of an ncu report: python scratch/ncu_lines.py rep.ncu-rep kernel_regex [top]"""
import collections, csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fpath, hdr, cur = None, None, None
inst, samp, text = collections.Counter(), collections.Counter(), {}
seen = set()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; ie = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); continue
    if hdr is None:
        continue
    if r[0] != "":
        cur = (fpath, int(r[0])); text[cur] = r[1].strip(); continue
    if len(r) > ie and r[2].startswith("0x") and r[2] not in seen:
        seen.add(r[2])
        try:
            inst[cur] += int(r[ie]); samp[cur] += int(r[si] or 0)
        except ValueError:
            pass
tot, ts = sum(inst.values()), sum(samp.values())
print("total warp instructions %d, samples %d" % (tot, ts))
for k, v in inst.most_common(top):
    print("%5.1f%% instr %5.1f%% samples  %s:%d  %s" % (100.0 * v / tot, 100.0 * samp[k] / max(1, ts), k[0], k[1], text[k][:110]))
