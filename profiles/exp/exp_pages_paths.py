"""~1 000-int lists (FastPFOR pages + VariableByte tails): batched kernels vs warp-per-list kernels, ms per call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mcp_b200
from mcp_b200 import native as N
ctx = mcp_b200.Context(0)
nl = 100_000
dv, do, total = ctx.generate_breach_lists(nl, 100_000, 0.01, 0x5EED0004)
cid = N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA
capb = 5 * total + 16 * nl
d_oo, d_out, d_back = ctx.dev_alloc(8 * (nl + 1)), ctx.dev_alloc(capb), ctx.dev_alloc(4 * total + 16)
for fast in (1, 0):
    ctx.set_option("fast_codec", fast)
    te, td = [], []
    for it in range(6):
        ctx.flush_l2(); ctx.sync(); ctx.timer_start()
        ctx.codec_encode_dev(cid, N.FRAME_APP, dv, do, nl, total, d_oo, d_out, capb)
        e = ctx.timer_stop()
        ctx.flush_l2(); ctx.sync(); ctx.timer_start()
        ctx.codec_decode_dev(cid, N.FRAME_APP, d_out, d_oo, nl, d_back, do)
        d = ctx.timer_stop()
        if it >= 2:
            te.append(e); td.append(d)
    print(f"fast_codec {fast}: encode {sum(te)/len(te):.3f} ms, decode {sum(td)/len(td):.3f} ms")
