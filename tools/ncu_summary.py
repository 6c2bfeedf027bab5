#!/usr/bin/env python
"""Summarises an .ncu-rep (raw page) into a markdown table of the metrics DESIGN.md / bench.py cite.
Usage: python tools/ncu_summary.py <report.ncu-rep> [kernel-name-substring ...] > profiles/xxx.md"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
filters = sys.argv[2:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = [
    ("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu (mufu) %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
]
idx = [(hdr.index(c), n) for c, n in cols if c in hdr]
print("| " + " | ".join(f"{n} [{units[i]}]" if units[i] else n for i, n in idx) + " |")
print("|" + "---|" * len(idx))
for r in data:
    name = r[hdr.index("Kernel Name")]
    if filters and not any(f in name for f in filters):
        continue
    cells = []
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            v = v.replace("void ", "").replace("sodt::<unnamed>::", "").split("(")[0][:48]
        else:
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
        cells.append(v)
    print("| " + " | ".join(cells) + " |")
