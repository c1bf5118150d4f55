import sys, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mcp_b200
from mcp_b200 import native as N
ctx = mcp_b200.Context(0)
P,Nn,F,alpha,seed=10000,100000,32,0.99,0x5EED0002
ctx.generate_exposures(P,F,seed,100.0,0.0); ctx.generate_scenarios(Nn,F,seed,1.0,0.05)
cid=N.CODEC_FASTPFOR256_VB|N.FLAG_DELTA
def bench(tag, steps=8):
    for _ in range(3): ctx.run(alpha,cid,N.FRAME_APP)
    ctx.profile(True); ctx.profile_reset()
    ms=[]
    for _ in range(steps):
        ctx.flush_l2(); ctx.sync(); ctx.timer_start(); ctx.run(alpha,cid,N.FRAME_APP); ms.append(ctx.timer_stop())
    pr=ctx.profile_get(); ctx.profile(False)
    print(tag, "step %.3f ms"%np.mean(ms), {k:round(v['ms']/steps,3) for k,v in pr.items() if v['launches']}, ctx.run_stats(), flush=True)
for ot in (4, 1, 2, 3, 5, 8):
    ctx.set_option("overlap_tiles", ot); bench("overlap_tiles=%d"%ot)
ctx.set_option("overlap_tiles", 1)
ctx.set_option("optimistic",0); bench("guaranteed"); ctx.set_option("optimistic",1)
