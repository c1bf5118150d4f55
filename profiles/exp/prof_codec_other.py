"""One encode + decode of the C4 lists under a codec id given by name (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mcp_b200
from mcp_b200 import native as N
ctx = mcp_b200.Context(0)
nl = 1_000_000
dv, do, total = ctx.generate_breach_lists(nl, 10_000, 0.01, 0x5EED0004)
name = sys.argv[1] if len(sys.argv) > 1 else "OPTPFD_VB"
cid = getattr(N, "CODEC_" + name) | (0 if name.startswith(("IBP", "SK_IBP", "XOR")) else N.FLAG_DELTA)
capb = 5 * total + 16 * nl
d_oo, d_out, d_back = ctx.dev_alloc(8 * (nl + 1)), ctx.dev_alloc(capb), ctx.dev_alloc(4 * total + 16)
for _ in range(2):
    ctx.codec_encode_dev(cid, N.FRAME_APP, dv, do, nl, total, d_oo, d_out, capb)
    ctx.codec_decode_dev(cid, N.FRAME_APP, d_out, d_oo, nl, d_back, do)
print("ok", ctx.codec_stats())
