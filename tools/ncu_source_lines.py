#!/usr/bin/env python
"""Warp-state samples of one kernel of an .ncu-rep (--set full --import-source on) summed per CUDA source line,
with the three largest stall reasons of each line.   python tools/ncu_source_lines.py rep.ncu-rep kernel_substring [min_pct]"""
import collections
import csv
import io
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur_file = cur_fn = hdr = None
agg = collections.defaultdict(collections.Counter)
src = {}
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]
    elif r[0] == "Function Name":
        cur_fn = r[1]
    elif r[0] == "Kernel Name":
        cur_fn = cur_file = None
    elif r[0] == "Line No":
        hdr = r
    elif cur_file and cur_fn and pat in cur_fn and hdr and len(r) == len(hdr) and r[0].isdigit():
        key = (cur_file.split("/")[-1], int(r[0]))
        src[key] = r[1]
        for h, v in zip(hdr, r):
            if h == "# Samples" or (h.startswith("stall_") and "Not Issued" not in h) or h == "Instructions Executed":
                try:
                    agg[key][h] += int(v)
                except ValueError:
                    pass
tot = sum(a["# Samples"] for a in agg.values())
print("total samples %d" % tot)
for key in sorted(agg):
    a = agg[key]
    if a["# Samples"] >= tot * min_pct / 100:
        top = sorted(((v, k) for k, v in a.items() if k.startswith("stall_")), reverse=True)[:3]
        print("%-22s %4d %5.1f%% exec=%9d  %-44s | %s" % (key[0], key[1], 100 * a["# Samples"] / tot, a["Instructions Executed"],
                                                        " ".join("%s:%d" % (k[6:], v) for v, k in top), src[key].strip()[:80]))
