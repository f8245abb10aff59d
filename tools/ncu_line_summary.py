#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS export with `nvdisasm -g` line info of the same cubin: stall samples and executed
warp instructions per source line.  usage: ncu_line_summary.py source.csv.gz cubin 'kernel-substring' [n_cells]"""
import csv, gzip, re, subprocess, sys, collections
fn, cubin, pat = sys.argv[1:4]
ncells = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
out = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
off2line = {}
cur, infunc = None, False
for l in out.splitlines():
    m = re.match(r'\s*\.text\.(\S+):', l)
    if m:
        infunc = pat in m.group(1)
        continue
    if not infunc: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m: off2line[int(m.group(1), 16)] = cur
f = gzip.open(fn, 'rt') if fn.endswith('.gz') else open(fn)
rows = list(csv.reader(f))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
base = int(rows[2][ix['Address']], 16)
agg = collections.defaultdict(lambda: [0, 0])
for r in rows[2:]:
    if len(r) < len(hdr): continue
    off = int(r[ix['Address']], 16) - base
    key = off2line.get(off)
    agg[key][0] += int(r[ix['# Samples']] or 0)
    agg[key][1] += int(r[ix['Instructions Executed']] or 0)
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
print('samples', ts, 'instructions', ti)
for k, v in sorted(agg.items(), key=lambda kv: (kv[0] or ('', 0))):
    if v[0] > 0.004 * ts or v[1] > 0.004 * ti:
        print('%-18s %5d  samples %5.1f%%  instr %5.1f%% (%7.1f/cell)' % ((k or ('?', 0))[0], (k or ('?', 0))[1], 100.0 * v[0] / ts, 100.0 * v[1] / ti, v[1] / ncells))
