#!/usr/bin/env python
"""Attribute executed SASS instructions of one kernel to the OUTERMOST source line of their inlining chain inside a given
file (nvdisasm -gi), joined with the executed counts of an ncu `--page source --csv` export of the same binary.
usage: sass_sections.py <src.csv> <nvdisasm -gi -c output> <function-substring> <file-substring> [top]"""
import collections
import csv
import re
import sys

src_csv, dis, func, want = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 60
rows = list(csv.reader(open(src_csv)))
start = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hdr = rows[start[0]]
body = rows[start[0] + 1:(start[1] - 1 if len(start) > 1 else len(rows))]
iE, iS, iSrc = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
chains, chain, on, fresh = [], [], False, True
for l in open(dis):
    if l.startswith('//-') and '.text.' in l:
        on = func in l
        continue
    if not on:
        continue
    m = re.match(r'\s*//## File "([^"]*)", line (\d+)', l)
    if m:
        if fresh:
            chain, fresh = [], False
        chain.append((m.group(1).split('/')[-1], int(m.group(2))))
        continue
    if re.match(r'\s*/\*[0-9a-f]{4,}\*/', l):
        chains.append(list(chain))
        fresh = True
print("sass instrs: ncu", len(body), "nvdisasm", len(chains))
nw = max(int(float(r[iE] or 0)) for r in body[:80])
ex, sm, cls = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
FP = {'FFMA', 'FMUL', 'FADD', 'MUFU', 'FSEL', 'FMNMX', 'FSETP'}
for k, r in enumerate(body[:len(chains)]):
    e = int(float(r[iE] or 0))
    key = None
    for f, ln in reversed(chains[k]):            # outermost frame first
        if want in f:
            key = (f, ln)
            break
    if key is None:
        key = chains[k][-1] if chains[k] else ('?', 0)
    ex[key] += e
    sm[key] += int(float(r[iS] or 0))
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[iSrc])
    op = m.group(2) if m else '?'
    cls[key]['fp' if op in FP else 'other'] += e
tot = sum(ex.values())
print(f"total executed per warp {tot / nw:.0f}")
for key, n in sorted(ex.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if n / nw >= 0.5:
        print(f"{key[0]}:{key[1]:<5d} {n / nw:7.1f} ({100 * n / tot:4.1f}%)  fp {cls[key]['fp'] / nw:6.1f} other {cls[key]['other'] / nw:6.1f}  samples {sm[key]:4d}")
