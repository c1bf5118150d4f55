#!/usr/bin/env python3
"""Evidence that survives without the binary: per kernel instantiation of libmcp.so, the counts of the SASS mnemonics
that matter (DMMA = FP64 tensor-core MMA, UBLKCP = cp.async.bulk / TMA bulk copy, SYNCS = mbarrier, LDGSTS = cp.async,
ATOM/RED, LDG/STG/LDS/STS, BAR) plus registers / shared memory from the ELF.

  python profiles/tools/sass_extract.py [libmcp.so] > profiles/r2/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "monte-carlo-portfolios_b200", "libmcp.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
cur = None
for ln in res.splitlines():
    m = re.match(r"\s*Function (\S+):", ln)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"REG:(\d+).*?SHARED:(\d+)", ln)
    if m and cur:
        usage[cur] = (int(m.group(1)), int(m.group(2)))
WATCH = ["DMMA", "UBLKCP", "SYNCS", "LDGSTS", "UTMALDG", "UTCMMA", "LDTM", "ATOM", "ATOMS", "RED", "LDG", "STG", "LDS", "STS", "BAR", "SHFL", "DFMA", "DADD", "DMUL"]
kern = None
counts = collections.OrderedDict()
arch = set()
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", ln)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and kern:
        op = m.group(1).split(".")[0]
        counts[kern]["_total"] += 1
        if op in WATCH:
            counts[kern][op] += 1
print(f"# {os.path.relpath(lib, ROOT)}: architectures {sorted(arch)}; {len(counts)} kernels")
print("# columns: SASS instructions | registers | static shared bytes | watched mnemonics (count)")
for k, c in counts.items():
    dem = subprocess.run(["cu++filt", k], capture_output=True, text=True).stdout.strip() or k
    reg, shm = usage.get(k, (None, None))
    watched = " ".join(f"{op}={c[op]}" for op in WATCH if c[op])
    print(f"{dem[:110]:110s} | {c['_total']:6d} | {reg} | {shm} | {watched}")
