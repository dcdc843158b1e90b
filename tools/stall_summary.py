#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv` (per-SASS-line warp-stall samples): totals per stall reason and the
top lines for the reasons that are not "pipe busy"."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
cols = ['stall_barrier', 'stall_dispatch', 'stall_long_sb', 'stall_math', 'stall_no_inst', 'stall_not_selected', 'stall_selected',
        'stall_short_sb', 'stall_wait', 'stall_branch_resolving', 'stall_mio', 'stall_lg']
def f(r, c):
    try: return float(r[ix[c]])
    except Exception: return 0.0
tot = {c: sum(f(r, c) for r in data) for c in cols}
T = sum(tot.values())
print({c: round(100 * v / T, 1) for c, v in tot.items()})
for c in ['stall_long_sb', 'stall_barrier', 'stall_short_sb', 'stall_wait', 'stall_dispatch', 'stall_no_inst']:
    print('==', c)
    for r in sorted(data, key=lambda r: -f(r, c))[:8]:
        print('  %6.0f  %s  %s' % (f(r, c), r[ix['Address']][-5:], r[ix['Source']][:90]))
