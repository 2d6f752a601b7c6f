#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel (name + grid): launches, total / average time, share."""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]; ci = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ci["Kernel Name"]]); name = re.sub(r"^void ", "", name)
    key = (name, r[ci["Grid Size"]], r[ci["Block Size"]])
    v = float(r[ci["Metric Value"]].replace(",", "")); unit = r[ci["Metric Unit"]]
    us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
    agg[key][0] += 1; agg[key][1] += us
tot = sum(v[1] for v in agg.values())
print("| kernel (grid x block) | launches | total ms | avg us | share |\n|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| %s (%s x %s) | %d | %.2f | %.1f | %.1f %% |" % (k[0], k[1].strip("()").replace(", 1, 1", ""), k[2].strip("()").replace(", 1, 1", ""), v[0], v[1] / 1e3, v[1] / v[0], 100 * v[1] / tot))
print("total %.2f ms" % (tot / 1e3))
