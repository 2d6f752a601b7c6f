#!/usr/bin/env python
"""Phase breakdown of a kernel from an `ncu --page source --print-source sass --csv` export + nvdisasm listing: segments between BAR.SYNC instructions;
barrier-stall samples (attributed by ncu to the instruction after the barrier) are moved back to the phase that was waited for.
usage: phase_profile.py src.csv all.sass <mangled kernel name>"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
lines, cur, on = [], ("?", 0), False
for ln in open(sys.argv[2]):
    if ln.startswith(".text."):
        on = ln.startswith(".text." + sys.argv[3] + ":"); continue
    if not on: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s*/\*[0-9a-f]+\*/", ln): lines.append(cur)
assert len(lines) == len(data), (len(lines), len(data))
IE, NS, SB = ci["Instructions Executed"], ci["# Samples"], ci["stall_barrier"]
idx = [i for i, r in enumerate(data) if "BAR.SYNC" in r[ci["Source"]]]
tot = sum(int(r[IE] or 0) for r in data); tots = sum(int(r[NS] or 0) for r in data)
segs, prev = [], 0
for b in idx + [len(data) - 1]:
    seg = data[prev:b + 1]
    segs.append([lines[prev], lines[b], sum(int(r[IE] or 0) for r in seg), sum(int(r[NS] or 0) for r in seg), sum(int(r[SB] or 0) for r in seg)]); prev = b + 1
print("phase (source lines) | instructions | own samples | + waited-for (barrier stalls of the next segment) | share of time")
for i, s in enumerate(segs):
    waited = segs[i + 1][4] if i + 1 < len(segs) else 0
    own = s[3] - s[4]
    print("%s:%d -> %d | %.1f %% | %.1f %% | %.1f %% | %.1f %%" % (s[0][0], s[0][1], s[1][1], 100 * s[2] / tot, 100 * own / tots, 100 * waited / tots, 100 * (own + waited) / tots))
