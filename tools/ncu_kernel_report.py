#!/usr/bin/env python
"""One captured launch of an .ncu-rep (--set full) as text: headline metrics, pipe utilisation, stall
reasons per issue, and the instruction-count share of each run of SASS lines with equal execution count.
Usage: python tools/ncu_kernel_report.py gpurun_out/prof.ncu-rep [launch index]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2 + idx]
want = ["Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for h, u, v in zip(hdr, units, r):
    if h in want or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(v or 0) >= 0.05):
        print("%-90s %-10s %s" % (h, u, v[:80]))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
blocks, names, cur, h = [], [], None, None
for row in srows:
    if row and row[0] == "Kernel Name":
        cur = []
        blocks.append(cur)
        names.append(row[1] if len(row) > 1 else "")
    elif row and row[0] == "Address":
        h = row
    elif cur is not None and len(row) > 10:
        cur.append(row)
# the SASS section is only trusted for single-kernel captures (ncu -k regex:<kernel>): block idx of the source page
if len({x[hdr.index("Kernel Name")] for x in rows[2:]}) != 1:
    print("\n(SASS runs are printed for captures of a single kernel only: use ncu -k regex:<name>)")
    sys.exit(0)
b = blocks[min(idx, len(blocks) - 1)]
iS, iE, iM = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
tot = sum(int(x[iE]) for x in b)
print("\nSASS runs (>= 0.3%% of %d executed warp instructions, %d SASS lines):" % (tot, len(b)))
run = None
for i, x in enumerate(b + [None]):
    e = int(x[iE]) if x else -1
    if run and x and abs(e - run[2]) <= 0.02 * max(e, run[2], 1):
        run[1], run[3], run[4] = i, run[3] + e, run[4] + int(x[iM])
        continue
    if run and run[3] > tot * 0.003:
        print("  %4d-%4d  n=%3d  exec/line=%9d  share=%5.1f%%  samples=%5d  %s" %
              (run[0], run[1], run[1] - run[0] + 1, run[2], 100.0 * run[3] / tot, run[4], b[run[0]][iS].strip()[:60]))
    if x:
        run = [i, i, e, e, int(x[iM])]
