#!/usr/bin/env python
"""Per-launch times of the n-th export in an `ncu --metrics gpu__time_duration.sum --csv` log:  launch_list.py <csv> [n]"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hdr = None
out = []
for r in rows:
    if r[0] == 'ID':
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    v = float(d['Metric Value'].replace(',', ''))
    u = d['Metric Unit']
    v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v
    name = d['Kernel Name']
    for prefix in ('tg::', 'tg_fast::'):
        if name.startswith(prefix):
            name = name[len(prefix):]
    out.append((name[:56].replace('tg::', ''), v, d.get('Grid Size', '')))
# exports start at CullRegionInitKernel; a pipelined export (host results) runs several slabs -- several brick kernels --
# behind one cull, a device-resident step exactly one: the n-th of those is listed
starts = [i for i, o in enumerate(out) if o[0].startswith('CullRegionInit')] + [len(out)]
steps = [(s, e) for s, e in zip(starts, starts[1:]) if sum(1 for o in out[s:e] if o[0].startswith('MeshBricksKernel')) == 1]
s, e = steps[min(which, len(steps) - 1)]
# drop what follows the export's last kernel (the L2 flush and the cull of nothing else belong to the next step)
while e > s and out[e - 1][0].startswith(('FillKernel', 'FmaChain')):
    e -= 1
total = sum(o[1] for o in out[s:e])
for o in out[s:e]:
    print("%-58s %9.1f us  %5.1f%%  grid %s" % (o[0], o[1], 100 * o[1] / total, o[2]))
print("total %.1f us" % total)
