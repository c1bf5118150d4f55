"""One C4 encode + decode of one codec (for ncu): argv = codec name, flat flag, lists."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mcp_b200
from mcp_b200 import native as N
name = sys.argv[1] if len(sys.argv) > 1 else "fp"
flat = int(sys.argv[2]) if len(sys.argv) > 2 else 1
nl = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
ntr = int(sys.argv[4]) if len(sys.argv) > 4 else 10_000
cid = {"fp": N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA, "bp": N.CODEC_BP32_VB | N.FLAG_DELTA, "vb": N.CODEC_VB | N.FLAG_DELTA}[name]
ctx = mcp_b200.Context(0)
ctx.set_option("flat_codec", flat)
dv, do, total = ctx.generate_breach_lists(nl, ntr, 0.01, 0x5EED0004)
capb = 5 * total + 16 * nl
d_oo, d_out, d_back = ctx.dev_alloc(8 * (nl + 1)), ctx.dev_alloc(capb), ctx.dev_alloc(4 * total + 16)
for it in range(3):
    ctx.flush_l2(); ctx.sync(); ctx.timer_start()
    nb = ctx.codec_encode_dev(cid, N.FRAME_APP, dv, do, nl, total, d_oo, d_out, capb)
    te = ctx.timer_stop()
    ctx.flush_l2(); ctx.sync(); ctx.timer_start()
    ctx.codec_decode_dev(cid, N.FRAME_APP, d_out, d_oo, nl, d_back, do)
    td = ctx.timer_stop()
print(name, flat, total, nb, te, td)
