"""Step time of the C2 shape for other confidence levels (tail sizes)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mcp_b200
from mcp_b200 import native as N
ctx = mcp_b200.Context(0)
P,Nn,F,seed=10000,100000,32,0x5EED0002
ctx.generate_exposures(P,F,seed,100.0,0.0); ctx.generate_scenarios(Nn,F,seed,1.0,0.05)
cid=N.CODEC_FASTPFOR256_VB|N.FLAG_DELTA
for alpha in (0.999, 0.99, 0.975, 0.95, 0.9, 0.5):
    for _ in range(2): ctx.run(alpha,cid,N.FRAME_APP)
    ctx.profile(True); ctx.profile_reset()
    ms=[]
    for _ in range(4):
        ctx.flush_l2(); ctx.sync(); ctx.timer_start(); ctx.run(alpha,cid,N.FRAME_APP); ms.append(ctx.timer_stop())
    pr=ctx.profile_get(); ctx.profile(False)
    print(alpha, ctx.run_sizes()["tail"], "step %.3f ms"%np.mean(ms), {k:round(v['ms']/4,3) for k,v in pr.items() if v['launches']}, ctx.run_stats(), flush=True)
