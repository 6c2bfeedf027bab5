#!/usr/bin/env python
"""Aggregates an ncu launch list (gpu__time_duration.sum per launch, --csv) per kernel name -> markdown."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for r in data:
    name = re.sub(r"\(.*", "", r[kn]).replace("void ", "")[:90]
    v = float(r[mv].replace(",", ""))
    v = v / 1e3 if r[mu] == "ns" else v * 1e3 if r[mu] == "ms" else v
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print(f"{len(data)} launches, {tot / 1e3:.2f} ms of device time\n")
print("| share | ms | launches | kernel |\n|---|---|---|---|")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if v / tot < 0.002:
        continue
    print(f"| {100 * v / tot:.1f} % | {v / 1e3:.2f} | {c} | `{k}` |")
