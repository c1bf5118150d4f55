"""C3 shape (125 000 x 10 000 x 64): step time vs the small_select option (64-thread CTAs vs one warp per row)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mcp_b200
from mcp_b200 import native as N
ctx = mcp_b200.Context(0)
P, Nn, F, alpha, seed = 125000, 10000, 64, 0.99, 0x5EED0003
ctx.generate_exposures(P, F, seed, 100.0, 0.0); ctx.generate_scenarios(Nn, F, seed, 1.0, 0.05)
cid = N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA
ref = None
for v in (0, 1, 0, 1):
    ctx.set_option("warp_select", v)
    for _ in range(3): ctx.run(alpha, cid, N.FRAME_APP)
    ctx.profile(True); ctx.profile_reset()
    ms = []
    for _ in range(5):
        ctx.flush_l2(); ctx.sync(); ctx.timer_start(); ctx.run(alpha, cid, N.FRAME_APP); ms.append(ctx.timer_stop())
    pr = ctx.profile_get(); ctx.profile(False)
    r = ctx.fetch()
    import hashlib
    h = hashlib.sha256(r["breach"].tobytes() + r["var_trial"].tobytes() + r["var_loss"].tobytes() + r["cvar"].tobytes()).hexdigest()[:12]
    print("warp_select", v, "step %.3f ms" % np.mean(ms), {k: round(x['ms'] / 5, 3) for k, x in pr.items() if x['launches']}, ctx.run_stats(), h, flush=True)
