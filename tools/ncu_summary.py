#!/usr/bin/env python
"""Summarise ncu output for profiles/: a launch list (--metrics gpu__time_duration.sum CSV)
into per-kernel totals/shares, and a --set full .ncu-rep into one line per captured launch
with the roofline-relevant metrics.  Usage:
  python tools/ncu_summary.py launches gpurun_out/launches.csv
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep [--traffic-json profiles/traffic.json]
"""
import collections
import csv
import json
import subprocess
import sys


def launches(path, skip=0, count=None):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr, agg = None, collections.OrderedDict()
    seen = 0
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        seen += 1
        if seen <= skip or (count is not None and seen > skip + count):
            continue
        d = dict(zip(hdr, r))
        v = float(d["Metric Value"].replace(",", ""))
        unit = d["Metric Unit"]
        v *= {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(unit, 1)
        a = agg.setdefault(d["Kernel Name"].split("(")[0][-48:], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("%-50s %6s %12s %12s %7s" % ("kernel", "n", "total_ms", "avg_us", "share"))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-50s %6d %12.3f %12.3f %6.1f%%" % (k, a[0], a[1] / 1e6, a[1] / a[0] / 1e3, 100 * a[1] / tot))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum"]


def full(path, traffic_json=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    traffic = {}
    for r in rows[2:]:
        name = r[ki].split("(")[0].split("::")[-1]
        print(name)
        vals = {}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                vals[w] = (r[i], units[i])
                print("    %-62s %16s %s" % (w, r[i], units[i]))
        try:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd = float(vals["dram__bytes_read.sum"][0].replace(",", "")) * scale[vals["dram__bytes_read.sum"][1]]
            wr = float(vals["dram__bytes_write.sum"][0].replace(",", "")) * scale[vals["dram__bytes_write.sum"][1]]
            t = traffic.setdefault(name, [])
            t.append(rd + wr)
        except Exception:
            pass
    if traffic_json:
        json.dump({k: max(v) for k, v in traffic.items()}, open(traffic_json, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
        count = int(sys.argv[sys.argv.index("--count") + 1]) if "--count" in sys.argv else None
        launches(sys.argv[2], skip, count)
    else:
        tj = sys.argv[sys.argv.index("--traffic-json") + 1] if "--traffic-json" in sys.argv else None
        full(sys.argv[2], tj)
