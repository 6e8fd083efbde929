#!/usr/bin/env python
"""Summarise an ncu report (CPU-side): key raw metrics, stall-reason totals, opcode mix and the top stalled SASS lines.
usage: ncu_summary.py <report.ncu-rep> [kernel-index]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum', 'sm__warps_active.avg.per_cycle_active']
for r in rows[2:]:
    print('---- launch')
    for w in want:
        for i, h in enumerate(hdr):
            if h == w:
                print(f'{w:70s} {units[i]:12s} {r[i]}')
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
start = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hdr = rows[start[0]]
body = rows[start[0] + 1:(start[1] - 1 if len(start) > 1 else len(rows))]
iE, iS, iSrc = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
s0, s1 = hdr.index('stall_barrier'), hdr.index('stall_wait') + 1
nw = max(int(float(r[iE] or 0)) for r in body[:80])
tot, ops = collections.Counter(), collections.Counter()
for r in body:
    for i in range(s0, s1):
        try:
            tot[hdr[i]] += int(float(r[i] or 0))
        except ValueError:
            pass
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[iSrc])
    ops[m.group(2) if m else '?'] += int(float(r[iE] or 0))
s = sum(tot.values())
print('---- stall samples (first launch):', s)
print('  '.join(f'{k[6:]} {100 * v / s:.1f}%' for k, v in tot.most_common(10)))
n = sum(ops.values())
print(f'---- executed warp-instructions per warp: {n / nw:.0f} (static {len(body)})')
print('  '.join(f'{k} {v / nw:.0f}' for k, v in ops.most_common(24)))
print('---- top stalled instructions')
for r in sorted(body, key=lambda r: -int(float(r[iS] or 0)))[:14]:
    reasons = {hdr[i][6:]: int(float(r[i] or 0)) for i in range(s0, s1) if r[i] and float(r[i]) > 0}
    print(r[0][-5:], r[iSrc][:64].ljust(64), r[iS], reasons)
