#!/usr/bin/env python
"""Prints the hottest SASS lines (by warp-stall samples) of `ncu --page source --csv` output.
Usage: ncu -i rep --page source --csv --kernel-name regex:X | python tools/ncu_hot.py [top] [section]"""
import csv
import sys

top_n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(sys.stdin))
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        sections.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
sec = sections[want]
hdr, data = sec["hdr"], sec["data"]
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(r[isamp] or 0) for r in data)
print(sec["name"][:100], "sections:", len(sections), "total samples", tot, "sass lines", len(data))
top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp] or 0))[:top_n]
for i in sorted(top):
    r = data[i]
    print(f"{i:5d} {int(r[isamp]):7d} {100 * int(r[isamp]) / tot:5.1f}%  x{r[iex]:>9}  {r[ia][:110]}")
