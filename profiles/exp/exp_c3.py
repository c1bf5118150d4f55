import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mcp_b200
from mcp_b200 import native as N
ctx = mcp_b200.Context(0)
P,Nn,F,alpha,seed=125000,10000,64,0.99,0x5EED0003
ctx.generate_exposures(P,F,seed,100.0,0.0); ctx.generate_scenarios(Nn,F,seed,1.0,0.05)
cid=N.CODEC_FASTPFOR256_VB|N.FLAG_DELTA
def bench(tag, steps=5):
    for _ in range(3): ctx.run(alpha,cid,N.FRAME_APP)
    ctx.profile(True); ctx.profile_reset()
    ms=[]
    for _ in range(steps):
        ctx.flush_l2(); ctx.sync(); ctx.timer_start(); ctx.run(alpha,cid,N.FRAME_APP); ms.append(ctx.timer_stop())
    pr=ctx.profile_get(); ctx.profile(False)
    print(tag, "step %.3f ms"%np.mean(ms), {k:round(v['ms']/steps,3) for k,v in pr.items() if v['launches']}, ctx.run_stats(), flush=True)
for fl in (2048, 1024, 768, 512, 256):
    ctx.set_option("dense_floor", fl); bench("dense_floor=%d"%fl)
