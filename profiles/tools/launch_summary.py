#!/usr/bin/env python3
"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`).

  python profiles/tools/launch_summary.py launches.csv "the command that was profiled" > summary.txt
"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 6 and r[0].isdigit()]
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("mcp::", "").strip()
    unit, val = r[-2], float(r[-1].replace(",", ""))
    us = val / 1e3 if unit in ("ns", "nsecond") else val * 1e3 if unit in ("ms", "msecond") else val
    tot[name][0] += 1
    tot[name][1] += us
allus = sum(v[1] for v in tot.values())
print(f"# ncu --metrics gpu__time_duration.sum --clock-control none --csv {sys.argv[2] if len(sys.argv) > 2 else ''}")
print("# (per-launch times under ncu are cold-cache and serialised: the SHARES are what carries over; all kernels are this repository's -- no library kernels)")
print(f"# {len(rows)} launches, {allus / 1e3:.1f} ms in total")
print("launches | total us | share | kernel")
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:6d} | {us:10.1f} | {100 * us / allus:5.1f}% | {name}")
