"""C2 step time vs the staging capacity of the sparse-step selection (option sel_slack adds entries to 2 t + 900)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mcp_b200
from mcp_b200 import native as N
P, Nn, F, alpha, seed = 10_000, 100_000, 32, 0.99, 0x5EED0002
ctx = mcp_b200.Context(0)
ctx.generate_exposures(P, F, seed, 100.0, 0.0)
ctx.generate_scenarios(Nn, F, seed, 1.0, 0.05)
cid = N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA
for v in (613, 0, -300, 300, 613, 0):
    ctx.set_option("sel_slack", v)
    ms = []
    ctx.profile(True)
    for it in range(9):
        if it == 3: ctx.profile_reset()
        ctx.flush_l2(); ctx.sync(); ctx.timer_start()
        ctx.run(alpha, cid, N.FRAME_APP)
        t = ctx.timer_stop()
        if it >= 3: ms.append(t)
    prof = ctx.profile_get()
    print("sel_slack", v, "cap", 2 * 1001 + 900 + v, f"{np.mean(ms):.3f} ms", {k: round(x["ms"] / 6, 3) for k, x in prof.items() if x["launches"]}, ctx.run_stats(), flush=True)
