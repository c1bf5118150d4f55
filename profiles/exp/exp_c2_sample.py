"""C2: size of the dense sample (option dense_mult: tail sizes in the first step) x staging slack of the sparse selection."""
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
ref = None
for mult, slack in ((4, 0), (3, 0), (3, 300), (2, 0), (2, 300), (2, 600), (4, 0)):
    ctx.set_option("dense_mult", mult)
    ctx.set_option("sel_slack", slack)
    ms = []
    ctx.profile(True)
    for it in range(8):
        if it == 3: ctx.profile_reset()
        ctx.flush_l2(); ctx.sync(); ctx.timer_start()
        ctx.run(alpha, cid, N.FRAME_APP)
        t = ctx.timer_stop()
        if it >= 3: ms.append(t)
    prof = ctx.profile_get()
    r = ctx.fetch()
    sig = (r["var_trial"].tobytes(), r["enc"].tobytes())
    if ref is None: ref = sig
    print("dense_mult", mult, "sel_slack", slack, f"{np.mean(ms):.3f} ms", {k: round(x["ms"] / 5, 3) for k, x in prof.items() if x["launches"] and k != "moments"},
          ctx.run_stats(), "same" if sig == ref else "DIFFERENT", flush=True)
