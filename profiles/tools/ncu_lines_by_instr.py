#!/usr/bin/env python3
"""Join an ncu source-page CSV (SASS rows with stall samples) with nvdisasm line info, aggregate by source line.

  cuobjdump -xelf all libmcp.so; nvdisasm --print-line-info risk_kernels.sm_100a.cubin > risk.sass
  ncu -i rep.ncu-rep --page source --csv > src.csv
  python ncu_lines.py src.csv risk.sass <mangled-kernel-substring> <part-index> [topN]
"""
import csv, io, re, sys, collections
src_csv, sass, kern, part = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
# --- line table from nvdisasm: offset -> (line, inlined-at chain ignored)
lines = {}
cur = None
infn = False
for ln in open(sass):
    if ln.startswith('.text.'):
        infn = kern in ln
        continue
    if not infn: continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
    if m:
        if 'inlined at' in ln and cur is not None and False: pass
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', ln)
    if m:
        lines[int(m.group(1), 16)] = (cur, m.group(2).strip())
txt = open(src_csv).read()
parts = re.split(r'\n(?="Kernel Name")', txt)
rows = list(csv.reader(io.StringIO(parts[part])))
h = rows[1]
si = h.index('# Samples'); ii = h.index('Instructions Executed')
stall_cols = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
base = int(rows[2][0], 16)
agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
tot_s = tot_i = 0
for r in rows[2:]:
    if len(r) <= si: continue
    off = int(r[0], 16) - base
    line = lines.get(off, (None, '?'))[0]
    s = float(r[si] or 0); n = float(r[ii] or 0)
    tot_s += s; tot_i += n
    a = agg[line]
    a[0] += s; a[1] += n
    for i in stall_cols:
        v = float(r[i] or 0)
        if v: a[2][h[i][6:]] += v
print(f"{rows[0][1][:60]}: samples {tot_s:.0f} instructions {tot_i:.0f}")
srcfile = {}
for line, (s, n, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    text = ''
    if line:
        f = '/root/repo/monte-carlo-portfolios_b200/csrc/' + line[0]
        if f not in srcfile:
            try: srcfile[f] = open(f).read().split('\n')
            except Exception: srcfile[f] = []
        L = srcfile[f]
        text = L[line[1] - 1].strip()[:90] if 0 < line[1] <= len(L) else ''
    print(f"{100*s/max(tot_s,1):5.1f}% smp {100*n/max(tot_i,1):5.1f}% ins  {str(line[1]) if line else '?':>5}  {text}   {dict((k,int(v)) for k,v in st.most_common(2))}")
