"""The local part of one C5 rank on one GPU, an eighth of the portfolios (12 500 x 1.25 M trials of 10 M, world 8): for launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mcp_b200
ctx = mcp_b200.Context(0)
P, Nl, F, alpha, seed, world = 12_500, 1_250_000, 32, 0.999, 0x5EED0005, 8
ctx.generate_exposures(P, F, seed, 100.0, 0.0)
ctx.generate_scenarios(Nl, F, seed, 1.0, 0.05, row0=0)
for _ in range(2):
    ctx.ts_local(alpha, Nl * world, 0, world)
ctx.sync()
print("ok")
