#!/usr/bin/env python
"""Join an ncu `--page source --csv` SASS listing with nvdisasm line info: executed warp-instructions and stall
samples per CUDA source line.  usage: sass_lines.py <src.csv> <nvdisasm -g -c output> <mangled-function-substring>"""
import collections
import csv
import re
import sys

src_csv, dis, func = sys.argv[1:4]
rows = list(csv.reader(open(src_csv)))
start = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hdr = rows[start[0]]
body = rows[start[0] + 1:(start[1] - 1 if len(start) > 1 else len(rows))]
iE, iS, iSrc = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
lines, cur, on = [], None, False
for l in open(dis):
    if l.startswith('//-') and '.text.' in l:
        on = func in l
        continue
    if not on:
        continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s*/\*[0-9a-f]{4,}\*/', l):
        lines.append(cur)
print("sass instrs: ncu", len(body), "nvdisasm", len(lines))
ex, sm = collections.Counter(), collections.Counter()
for k, r in enumerate(body[:len(lines)]):
    try:
        ex[lines[k]] += int(float(r[iE] or 0)); sm[lines[k]] += int(float(r[iS] or 0))
    except ValueError:
        pass
tot, tots = sum(ex.values()), sum(sm.values())
byfile = collections.Counter()
for (f, ln), n in ex.items():
    byfile[f] += n
print("per file:", [(f, round(100 * n / tot, 1)) for f, n in byfile.most_common(8)])
nwarps = max(int(float(r[iE] or 0)) for r in body[:50])
print(f"total executed {tot} (= {tot / nwarps:.0f} per warp), samples {tots}")
for (f, ln), n in ex.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 45):
    print(f"{f}:{ln:<5d} exec/warp {n / nwarps:7.1f} ({100 * n / tot:4.1f}%)  samples {sm[(f, ln)]:4d}")
