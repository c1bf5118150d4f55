#!/usr/bin/env python3
"""Summarise `ncu -i X.ncu-rep --page source --csv` output: top source lines by stall samples / instructions.
usage: ncu -i rep --page source --csv > src.csv ; python ncu_src.py src.csv [kernel-index] [topN]"""
import csv, io, re, sys
txt = open(sys.argv[1]).read()
parts = re.split(r'\n(?="Kernel Name")', txt)
sel = int(sys.argv[2]) if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
for pi, part in enumerate(parts):
    rows = list(csv.reader(io.StringIO(part)))
    name = rows[0][1][:70]
    h = rows[1]
    first = h[0]
    print(f"== part {pi}: {name}  view={first} rows={len(rows)-2}")
    if sel is not None and pi != sel:
        continue
    si = h.index('# Samples'); ii = h.index('Instructions Executed'); src = h.index('Source')
    tot_s = sum(float(r[si] or 0) for r in rows[2:] if len(r) > si)
    tot_i = sum(float(r[ii] or 0) for r in rows[2:] if len(r) > ii)
    stall_cols = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
    data = []
    for r in rows[2:]:
        if len(r) <= si: continue
        s = float(r[si] or 0); n = float(r[ii] or 0)
        st = sorted(((float(r[i] or 0), h[i][6:]) for i in stall_cols), reverse=True)[:2]
        data.append((s, n, r[0], r[src][:110], st))
    print(f"   total samples {tot_s:.0f}, instructions {tot_i:.0f}")
    for s, n, a, t, st in sorted(data, reverse=True)[:top]:
        print(f"   {100*s/max(tot_s,1):5.1f}% smp {100*n/max(tot_i,1):5.1f}% ins  {a:>6} {t}   {[(x[1], int(x[0])) for x in st]}")
