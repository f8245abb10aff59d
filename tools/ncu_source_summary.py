#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export (SASS view): executed warp instructions by opcode, stall samples by
reason, and the hottest SASS lines.  usage: ncu_source_summary.py file.csv[.gz] [n_cells]"""
import csv, gzip, sys, collections
fn = sys.argv[1]
ncells = float(sys.argv[2]) if len(sys.argv) > 2 else None
f = gzip.open(fn, 'rt') if fn.endswith('.gz') else open(fn)
rows = list(csv.reader(f))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); samples = collections.Counter(); tot = 0
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
stalls = collections.Counter()
lines = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix['Source']].strip()
    op = src.split()[0] if src else '?'
    if op.startswith('@'): op = src.split()[1]
    op = op.split('.')[0]
    n = int(r[ix['Instructions Executed']] or 0)
    s = int(r[ix['# Samples']] or 0)
    ops[op] += n; samples[op] += s; tot += n
    for c in stall_cols: stalls[c] += int(r[ix[c]] or 0)
    lines.append((s, n, src, r[ix['Address']]))
print('total warp instructions', tot, ('= %.0f per cell' % (tot / ncells)) if ncells else '')
ts = sum(samples.values())
for op, n in ops.most_common(22):
    print('%-10s %12d %6.1f%%  samples %5.1f%%' % (op, n, 100.0 * n / tot, 100.0 * samples[op] / max(ts, 1)), ('%8.1f/cell' % (n / ncells)) if ncells else '')
print('stall samples:', ', '.join('%s %.1f%%' % (k.replace('stall_', ''), 100.0 * v / max(sum(stalls.values()), 1)) for k, v in stalls.most_common(9)))
print('hottest lines:')
for s, n, src, a in sorted(lines, reverse=True)[:14]:
    print('  %5d samples  %10d exec  %s' % (s, n, src[:90]))
