"""C2 valuation time vs the work-item size (option item_trials -> sub-chunks per step -> CTAs per launch)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mcp_b200
from mcp_b200 import native as N
P, Nn, F, alpha, seed = 10_000, 100_000, 32, 0.99, 0x5EED0002
if len(sys.argv) > 1 and sys.argv[1] == "c3":
    P, Nn, F, alpha, seed = 125_000, 10_000, 64, 0.99, 0x5EED0003
ctx = mcp_b200.Context(0)
ctx.generate_exposures(P, F, seed, 100.0, 0.0)
ctx.generate_scenarios(Nn, F, seed, 1.0, 0.05)
cid = N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA
for v in (2048, 2200, 2500, 2048):
    ctx.set_option("item_trials", v)
    ms = []
    ctx.profile(True)
    for it in range(8):
        if it == 3: ctx.profile_reset()
        ctx.flush_l2(); ctx.sync(); ctx.timer_start()
        ctx.run(alpha, cid, N.FRAME_APP)
        t = ctx.timer_stop()
        if it >= 3: ms.append(t)
    prof = ctx.profile_get()
    print("item_trials", v, f"{np.mean(ms):.3f} ms", {k: round(x["ms"] / 5, 3) for k, x in prof.items() if x["launches"]}, flush=True)
