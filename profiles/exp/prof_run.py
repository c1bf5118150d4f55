"""Two passes of the C2 step (for ncu captures: run plain first, then under ncu with -k / -s / -c)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mcp_b200
from mcp_b200 import native as N
ctx = mcp_b200.Context(0)
P, Nn, F, alpha, seed = 10000, 100000, 32, 0.99, 0x5EED0002
for a in sys.argv[1:]:
    k, v = a.split("=")
    if k == "P": P = int(v)
    elif k == "N": Nn = int(v)
    elif k == "F": F = int(v)
    else: ctx.set_option(k, int(v))
ctx.generate_exposures(P, F, seed, 100.0, 0.0)
ctx.generate_scenarios(Nn, F, seed, 1.0, 0.05)
for _ in range(2):
    ctx.run(alpha, N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA, N.FRAME_APP)
ctx.sync()
print("ok", ctx.run_stats())
