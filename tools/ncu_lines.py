#!/usr/bin/env python
"""Join an `ncu --page source --print-source sass --csv` export with `nvdisasm -g` line info and print the share of
executed warp instructions per source line (innermost inlined location).  Usage:
    ncu -i prof.ncu-rep --page source --csv --print-source sass > sass.csv
    cuobjdump -xelf all libmpcmmd.so ; nvdisasm -g -c x.cubin > all.sass
    python tools/ncu_lines.py sass.csv all.sass '<mangled kernel name>' [top_n]
"""
import csv
import re
import sys
from collections import Counter


def main():
    sass_csv, disasm, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
    top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(sass_csv)))
    hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
    data = []
    for r in rows[2:]:
        if len(r) != len(hdr) or r[0] == "Address":
            break
        data.append(r)
    lines, cur, on = [], ("?", 0), False
    for ln in open(disasm):
        if ln.startswith(".text."):
            on = ln.startswith(".text." + kernel + ":")
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]+\*/", ln):
            lines.append(cur)
    assert len(lines) == len(data), (len(lines), len(data))
    IE, TI, NS = ci["Instructions Executed"], ci["Thread Instructions Executed"], ci["# Samples"]
    tot = sum(int(r[IE] or 0) for r in data); tots = sum(int(r[NS] or 0) for r in data)
    by, bys = Counter(), Counter()
    for loc, r in zip(lines, data):
        by[loc] += int(r[IE] or 0); bys[loc] += int(r[NS] or 0)
    print(f"total warp instructions {tot}, stall samples {tots}, sass lines {len(data)}")
    for loc, v in by.most_common(top_n):
        print(f"{loc[0]}:{loc[1]:<5d} inst {v / tot * 100:6.2f}%  samples {bys[loc] / max(tots, 1) * 100:6.2f}%")


if __name__ == "__main__":
    main()
