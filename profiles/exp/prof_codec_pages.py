"""One encode + decode of the ~1000-int-list workload (FastPFOR pages + VariableByte tails), for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mcp_b200
from mcp_b200 import native as N
ctx = mcp_b200.Context(0)
nl = 100_000
dv, do, total = ctx.generate_breach_lists(nl, 100_000, 0.01, 0x5EED0004)
cid = (N.CODEC_FASTPFOR256_VB if len(sys.argv) < 2 or sys.argv[1] == "fp" else N.CODEC_BP32_VB) | N.FLAG_DELTA
capb = 5 * total + 16 * nl
d_oo, d_out, d_back = ctx.dev_alloc(8 * (nl + 1)), ctx.dev_alloc(capb), ctx.dev_alloc(4 * total + 16)
for _ in range(3):
    ctx.codec_encode_dev(cid, N.FRAME_APP, dv, do, nl, total, d_oo, d_out, capb)
    ctx.codec_decode_dev(cid, N.FRAME_APP, d_out, d_oo, nl, d_back, do)
print("ok", ctx.codec_stats())
